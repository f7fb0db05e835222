"""DETRsegm: the detector plus the mask branch, behind the reference's interface (models/segmentation.py:17-168).

`DETRsegm(detr, mask_head="smallconv", freeze_detr=...)` keeps the reference's attribute / state-dict names
(`detr.*`, `bbox_attention.{q_linear,k_linear}.*`, `mask_head.{lay1-5,gn1-5,out_lay,adapter1-3}.*`) and the two-phase
forward; phase B adds `pred_masks` [B, Q, H/4, W/4].  The arithmetic of the branch is toist_b200/maskhead.py.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from ..maskhead import MASKHEAD
from ..runtime import Call, GraphCache, ShadowBank, Stage, run_stage
from ..util.misc import NestedTensor


class MHAttentionMap(nn.Module):
    """Parameter container of the 2-D attention module that only returns the per-head softmax over pixels."""

    def __init__(self, query_dim, hidden_dim, num_heads, dropout=0, bias=True):
        super().__init__()
        if dropout != 0 or not bias:
            raise NotImplementedError("MHAttentionMap is built with dropout=0, bias=True in the reference")
        self.num_heads = num_heads
        self.hidden_dim = hidden_dim
        self.q_linear = nn.Linear(query_dim, hidden_dim, bias=bias)
        self.k_linear = nn.Linear(query_dim, hidden_dim, bias=bias)
        nn.init.zeros_(self.k_linear.bias)
        nn.init.zeros_(self.q_linear.bias)
        nn.init.xavier_uniform_(self.k_linear.weight)
        nn.init.xavier_uniform_(self.q_linear.weight)
        self.normalize_fact = float(hidden_dim / self.num_heads) ** -0.5


class MaskHeadSmallConv(nn.Module):
    """Parameter container: 5 x (3x3 conv + GroupNorm(8)) with three 1x1 FPN adapters and a 3x3 conv to one channel."""

    def __init__(self, dim, fpn_dims, context_dim):
        super().__init__()
        inter = [dim, context_dim // 2, context_dim // 4, context_dim // 8, context_dim // 16, context_dim // 64]
        self.lay1 = nn.Conv2d(dim, dim, 3, padding=1)
        self.gn1 = nn.GroupNorm(8, dim)
        self.lay2 = nn.Conv2d(dim, inter[1], 3, padding=1)
        self.gn2 = nn.GroupNorm(8, inter[1])
        self.lay3 = nn.Conv2d(inter[1], inter[2], 3, padding=1)
        self.gn3 = nn.GroupNorm(8, inter[2])
        self.lay4 = nn.Conv2d(inter[2], inter[3], 3, padding=1)
        self.gn4 = nn.GroupNorm(8, inter[3])
        self.lay5 = nn.Conv2d(inter[3], inter[4], 3, padding=1)
        self.gn5 = nn.GroupNorm(8, inter[4])
        self.out_lay = nn.Conv2d(inter[4], 1, 3, padding=1)
        self.dim = dim
        self.adapter1 = nn.Conv2d(fpn_dims[0], inter[1], 1)
        self.adapter2 = nn.Conv2d(fpn_dims[1], inter[2], 1)
        self.adapter3 = nn.Conv2d(fpn_dims[2], inter[3], 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, a=1)
                nn.init.constant_(m.bias, 0)


class _SegRuntime:
    def __init__(self):
        self.bank = ShadowBank()
        self.stage: Optional[Stage] = None
        self.graphs: Optional[GraphCache] = None
        self.dirty = True
        self.steps = 0

    def __deepcopy__(self, memo):
        return _SegRuntime()


class DETRsegm(nn.Module):
    def __init__(self, detr, mask_head="smallconv", freeze_detr=False):
        super().__init__()
        self.detr = detr
        if freeze_detr:
            for p in self.parameters():
                p.requires_grad_(False)
        hidden_dim, nheads = detr.transformer.d_model, detr.transformer.nhead
        self.bbox_attention = MHAttentionMap(hidden_dim, hidden_dim, nheads, dropout=0)
        if mask_head != "smallconv":
            raise RuntimeError(f"Unknown mask model {mask_head}")
        if not detr.backbone[0].return_interm_layers:
            raise RuntimeError("DETRsegm needs the backbone's intermediate layers (args.masks, main.py:297-298)")
        self.mask_head = MaskHeadSmallConv(hidden_dim + nheads, [1024, 512, 256], hidden_dim)
        self._rt = _SegRuntime()

    def _apply(self, fn, *args, **kwargs):
        self._rt.dirty = True
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._rt.dirty = True
        self.detr._rt.dirty = True
        return super().load_state_dict(*args, **kwargs)

    def enable_cuda_graphs(self, on: bool = True) -> "DETRsegm":
        self.detr.enable_cuda_graphs(on)
        self._rt.graphs = GraphCache() if on else None
        return self

    def enable_direct_grads(self, on: bool = True) -> "DETRsegm":
        self.detr.enable_direct_grads(on)
        return self

    def _refresh(self) -> None:
        rt = self._rt
        if rt.stage is None:
            names = [n for n, _ in self.named_parameters() if n.startswith(("bbox_attention.", "mask_head."))]
            rt.stage = Stage(self, "maskhead", names, nheads=self.detr.transformer.nhead)
        rt.steps += 1
        dirty = rt.dirty or rt.steps % 64 == 0
        rt.dirty = False
        if dirty:
            rt.stage.invalidate()
        sig = rt.bank._sig
        rt.bank.ensure(self, None, None, dirty, only=("bbox_attention.", "mask_head."))
        if rt.graphs is not None and sig is not None and sig != rt.bank._sig:
            rt.graphs.clear()

    def forward(self, samples: NestedTensor, captions, encode_and_save=True, memory_cache=None):
        if not isinstance(samples, NestedTensor):
            samples = NestedTensor.from_tensor_list(samples)
        if encode_and_save:
            assert memory_cache is None
            return self.detr.encode(samples, captions, want_features=True)
        assert memory_cache is not None
        out = self.detr.decode(memory_cache, want_hs=True)
        self._refresh()
        rt = self._rt
        hs = out.pop("_b200_hs")
        feats = memory_cache["_b200_feats"]  # NHWC bf16 (c2, c3, c4, c5)
        mem = memory_cache["img_memory"]
        save = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        call = Call(rt.stage, rt.bank.w, save, graphs=rt.graphs, n_dec_layers=int(hs.shape[0]), seq_len=int(mem.shape[0]))
        call.grad_sync = getattr(rt, "grad_sync", None)
        if self.detr._rt.direct and rt.stage.params:
            if self.detr._rt.anchor is None:
                self.detr._rt.anchor = torch.zeros(1, device=rt.stage.params[0].device, requires_grad=True)
            call.anchor = self.detr._rt.anchor
        pred = run_stage(MASKHEAD, call, hs, mem, memory_cache["_b200_src_proj"], feats[2], feats[1], feats[0],
                         memory_cache["_b200_small_mask"])[0]
        out["pred_masks"] = pred
        return out
