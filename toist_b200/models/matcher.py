"""HungarianMatcher on the device (reference models/matcher.py:15-99).

The block-diagonal matching cost of every (decoder layer, image) pair comes from one `toist_match_cost` launch and the
assignments from one `toist_lsap_device` launch (float64 shortest augmenting paths, scipy's tie rules), so the
criterion never synchronises with the host.  `HungarianMatcher.forward` keeps the reference's public contract
(CPU int64 index pairs, ValueError on an invalid cost matrix) for callers that want the indices themselves.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Tuple

import torch
from torch import nn

from .. import kernels as K
from ..util.misc import h2d


class PackedTargets(NamedTuple):
    boxes: torch.Tensor   # [B, t_max, 4] fp32 cxcywh, zero padded
    count: torch.Tensor   # [B] int32
    posmap: torch.Tensor  # [B, t_max, C] fp32
    counts: Tuple[int, ...]
    t_max: int


def pack_targets(targets: Sequence[dict], positive_map: torch.Tensor, device, t_max: Optional[int] = None) -> PackedTargets:
    """Pads the per-image target lists to [B, t_max, ...] so that one launch covers the whole batch
    (the reference concatenates them instead: models/matcher.py:66, models/mdetr.py:497-505)."""
    counts = tuple(int(t["boxes"].shape[0]) for t in targets)
    B = len(targets)
    tm = max(max(counts) if counts else 0, 1)
    if t_max is not None:
        assert t_max >= tm
        tm = t_max
    C = positive_map.shape[-1]
    assert positive_map.shape[0] == sum(counts), "positive_map rows must equal the number of target boxes"
    tb = torch.zeros((B * tm, 4), dtype=torch.float32, device=device)
    pp = torch.zeros((B * tm, C), dtype=torch.float32, device=device)
    if sum(counts):
        rows = [b * tm + t for b, n in enumerate(counts) for t in range(n)]
        idx = h2d(torch.tensor(rows, dtype=torch.int64), device)
        cat = h2d(torch.cat([t["boxes"].reshape(-1, 4) for t in targets]), device, torch.float32)
        tb.index_copy_(0, idx, cat)
        pp.index_copy_(0, idx, h2d(positive_map, device, torch.float32))
    cnt = h2d(torch.tensor(counts, dtype=torch.int32), device)
    return PackedTargets(tb.view(B, tm, 4), cnt, pp.view(B, tm, C), counts, tm)


def match_layers(logits: torch.Tensor, boxes: torch.Tensor, pt: PackedTargets, w_class: float, w_bbox: float,
                 w_giou: float):
    """logits [L,B,Q,C], boxes [L,B,Q,4] fp32 -> (match_q int32 [L,B,t_max], flags int32 [1], cost)."""
    cost = K.match_cost(logits, boxes, pt.boxes, pt.count, pt.posmap, w_class, w_bbox, w_giou)
    flags = torch.zeros(1, dtype=torch.int32, device=logits.device)
    match_q = K.lsap_device(cost, pt.count, flags)
    return match_q, flags, cost


def indices_from_match(match_q_cpu: torch.Tensor, counts: Sequence[int]) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """[B, t_max] query-per-target -> the reference's per-image (query_idx ascending, target_idx) int64 pairs."""
    out = []
    for b, n in enumerate(counts):
        q = match_q_cpu[b, :n].to(torch.int64)
        t = torch.arange(n, dtype=torch.int64)
        keep = q >= 0
        q, t = q[keep], t[keep]
        order = torch.argsort(q)
        out.append((q[order], t[order]))
    return out


class HungarianMatcher(nn.Module):
    """1-to-1 assignment between predictions and targets minimising
    cost_bbox * L1 + cost_class * (-prob . positive_map) + cost_giou * (-GIoU)."""

    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        self.norm = nn.Softmax(-1)
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"

    @torch.no_grad()
    def forward(self, outputs, targets, positive_map):
        logits = outputs["pred_logits"].detach().float().contiguous()
        boxes = outputs["pred_boxes"].detach().float().contiguous()
        bs, nq = logits.shape[:2]
        assert sum(len(v["boxes"]) for v in targets) == len(positive_map)
        pt = pack_targets(targets, positive_map, logits.device)
        match_q, flags, _ = match_layers(logits[None], boxes[None], pt, float(self.cost_class), float(self.cost_bbox),
                                         float(self.cost_giou))
        if int(flags.item()) != 0:  # scipy.optimize.linear_sum_assignment raises ValueError here
            raise ValueError("matrix contains invalid numeric entries")
        return indices_from_match(match_q[0].cpu(), pt.counts)


def build_matcher(args):
    if args.set_loss != "hungarian":
        raise ValueError(f"Only hungarian accepted, got {args.set_loss}")
    return HungarianMatcher(cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox,
                            cost_giou=args.set_cost_giou)
