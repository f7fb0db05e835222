"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch fp32, functional, state-dict driven) of the reference's
MDETR/TOIST forward + criterion.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this package; nothing under toist_b200/ does.

Parity status: **pinned by generated goldens** — the reference ships no tests or known-answer vectors (SURVEY.md §4,
§8c).  tests/test_oracle_vs_reference.py checks every function here against the unmodified reference modules imported
through oracle/shims.py (this container only), and tools/make_golden.py freezes reference outputs under tests/golden/.

All `file:line` citations are into /root/reference.  Dropout is not restated: the oracle is defined for
`model.eval()` / `--dropout 0` (dropout RNG streams cannot match across implementations).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# ===================================================================================================== box ops


def box_cxcywh_to_xyxy(b: Tensor) -> Tensor:
    """util/box_ops.py:11-14"""
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def pairwise_giou(a: Tensor, b: Tensor) -> Tensor:
    """util/box_ops.py:24-61 (box_iou + generalized_box_iou), xyxy inputs, returns [len(a), len(b)]."""
    assert bool((a[:, 2:] >= a[:, :2]).all()) and bool((b[:, 2:] >= b[:, :2]).all())
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = area_a[:, None] + area_b[None, :] - inter
    iou = inter / union
    lt2 = torch.min(a[:, None, :2], b[None, :, :2])
    rb2 = torch.max(a[:, None, 2:], b[None, :, 2:])
    wh2 = (rb2 - lt2).clamp(min=0)
    hull = wh2[..., 0] * wh2[..., 1]
    return iou - (hull - union) / hull


# ===================================================================================================== LSAP


def lsap(cost: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Rectangular linear sum assignment, restating the algorithm scipy.optimize.linear_sum_assignment documents
    (Crouse 2016, "On implementing 2D rectangular assignment algorithms": shortest augmenting paths with dual
    variables u, v; float64; the matrix is transposed when it has more rows than columns; result sorted by row).
    scipy is a compiled third-party dependency of the reference (requirements.txt:67, called at
    models/matcher.py:85, models/mdetr.py:100,539); tests pin this restatement against the installed scipy.
    """
    c = np.asarray(cost, dtype=np.float64)
    if c.ndim != 2:
        raise ValueError("expected a matrix")
    if c.size and (np.isnan(c).any() or np.isneginf(c).any()):
        raise ValueError("matrix contains invalid numeric entries")
    transposed = c.shape[0] > c.shape[1]
    if transposed:
        c = c.T
    nr, nc = c.shape
    if nr == 0 or nc == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    u = np.zeros(nr)
    v = np.zeros(nc)
    col4row = np.full(nr, -1, dtype=np.int64)
    row4col = np.full(nc, -1, dtype=np.int64)
    for cur_row in range(nr):
        # shortest augmenting path from cur_row (Dijkstra over reduced costs)
        shortest = np.full(nc, np.inf)
        path = np.full(nc, -1, dtype=np.int64)
        SR = np.zeros(nr, dtype=bool)
        SC = np.zeros(nc, dtype=bool)
        remaining = list(range(nc - 1, -1, -1))  # scipy fills remaining[it] = nc - it - 1
        min_val = 0.0
        i = cur_row
        sink = -1
        while sink == -1:
            index = -1
            lowest = np.inf
            SR[i] = True
            for it, j in enumerate(remaining):
                r = min_val + c[i, j] - u[i] - v[j]
                if r < shortest[j]:
                    path[j] = i
                    shortest[j] = r
                # ties are broken in favour of an unassigned column (a sink), as scipy does
                if shortest[j] < lowest or (shortest[j] == lowest and row4col[j] == -1):
                    lowest = shortest[j]
                    index = it
            min_val = lowest
            if min_val == np.inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            remaining[index] = remaining[-1]
            remaining.pop()
        # dual update
        u[cur_row] += min_val
        for r_i in range(nr):
            if SR[r_i] and r_i != cur_row:
                u[r_i] += min_val - shortest[col4row[r_i]]
        for j in range(nc):
            if SC[j]:
                v[j] -= min_val - shortest[j]
        # augment
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur_row:
                break
    if transposed:
        order = np.argsort(col4row, kind="stable")
        return col4row[order].astype(np.int64), order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), col4row.astype(np.int64)


# ===================================================================================================== matcher


def matcher_cost(pred_logits: Tensor, pred_boxes: Tensor, tgt_boxes: Tensor, positive_map: Tensor,
                 w_class: float = 1.0, w_bbox: float = 5.0, w_giou: float = 2.0) -> Tensor:
    """models/matcher.py:60-82 — returns C as [B, Q, sum(T_i)] fp32 (operation order kept)."""
    bs, nq = pred_logits.shape[:2]
    prob = pred_logits.flatten(0, 1).softmax(-1)
    boxes = pred_boxes.flatten(0, 1)
    assert len(tgt_boxes) == len(positive_map)
    cost_class = -(prob.unsqueeze(1) * positive_map.unsqueeze(0)).sum(-1)
    cost_bbox = torch.cdist(boxes, tgt_boxes, p=1)
    cost_giou = -pairwise_giou(box_cxcywh_to_xyxy(boxes), box_cxcywh_to_xyxy(tgt_boxes))
    c = w_bbox * cost_bbox + w_class * cost_class + w_giou * cost_giou
    return c.view(bs, nq, -1)


def hungarian_match(pred_logits: Tensor, pred_boxes: Tensor, targets: Sequence[dict], positive_map: Tensor,
                    w_class: float = 1.0, w_bbox: float = 5.0, w_giou: float = 2.0) -> List[Tuple[Tensor, Tensor]]:
    """models/matcher.py:39-87 — per-image assignment on the block-diagonal of the cost matrix."""
    with torch.no_grad():
        tgt_boxes = torch.cat([t["boxes"] for t in targets])
        c = matcher_cost(pred_logits, pred_boxes, tgt_boxes, positive_map, w_class, w_bbox, w_giou).cpu()
        sizes = [len(t["boxes"]) for t in targets]
        out = []
        for i, blk in enumerate(c.split(sizes, -1)):
            r, cidx = lsap(blk[i].numpy())
            out.append((torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(cidx, dtype=torch.int64)))
        return out


# ===================================================================================================== backbone


def frozen_bn(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """models/backbone.py:48-58 (eps 1e-5, scale = w * rsqrt(var + eps), bias = b - mean * scale)."""
    scale = sd[prefix + "weight"] * (sd[prefix + "running_var"] + 1e-5).rsqrt()
    shift = sd[prefix + "bias"] - sd[prefix + "running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


RESNET_BLOCKS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


def resnet_body(x: Tensor, sd: SD, prefix: str, arch: str, dilation: bool = False) -> List[Tensor]:
    """torchvision resnet50/101 (Bottleneck v1.5: stride on the 3x3) with FrozenBatchNorm, as wired by
    models/backbone.py:83-91; returns the outputs of layer1..layer4."""
    assert not dilation, "dilation (DC5) is out of scope (main.py:99-103 default False)"
    x = F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3)
    x = F.relu(frozen_bn(x, sd, prefix + "bn1."))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for li, nblocks in enumerate(RESNET_BLOCKS[arch], start=1):
        for bi in range(nblocks):
            p = f"{prefix}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            idt = x
            y = F.relu(frozen_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1."))
            y = F.relu(frozen_bn(F.conv2d(y, sd[p + "conv2.weight"], stride=stride, padding=1), sd, p + "bn2."))
            y = frozen_bn(F.conv2d(y, sd[p + "conv3.weight"]), sd, p + "bn3.")
            if (p + "downsample.0.weight") in sd:
                idt = frozen_bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), sd, p + "downsample.1.")
            x = F.relu(y + idt)
        outs.append(x)
    return outs


def downsample_mask(mask: Tensor, size: Tuple[int, int]) -> Tensor:
    """models/backbone.py:78 — nearest resize of the bool padding mask."""
    return F.interpolate(mask[None].float(), size=size).bool()[0]


def position_sine(mask: Tensor, num_pos_feats: int = 128, temperature: float = 10000.0) -> Tensor:
    """models/position_encoding.py:30-49 with normalize=True, scale 2*pi, eps 1e-6. mask [B,H,W] -> [B,2F,H,W]."""
    not_mask = ~mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=mask.device)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


# ===================================================================================================== text encoder


def roberta_encode(input_ids: Tensor, attention_mask: Tensor, sd: SD, prefix: str, num_layers: int = 12,
                   num_heads: int = 12, eps: float = 1e-12, pad_id: int = 1) -> Tensor:
    """transformers RobertaModel (BERT-style post-LN encoder, erf-GELU) as called at models/transformer.py:130.
    `eps` is RobertaConfig.layer_norm_eps: 1e-12 for the random-init config (transformer.py:61), 1e-5 for the
    roberta-base checkpoint.  Returns last_hidden_state [B, L, 768]."""
    e = prefix + "embeddings."
    nonpad = input_ids.ne(pad_id).int()
    pos_ids = (torch.cumsum(nonpad, dim=1) * nonpad).long() + pad_id
    # both tables are nn.Embedding(..., padding_idx=pad_token_id) in transformers' RobertaEmbeddings: the rows of
    # padded positions receive no gradient
    x = F.embedding(input_ids, sd[e + "word_embeddings.weight"], padding_idx=pad_id) + \
        F.embedding(pos_ids, sd[e + "position_embeddings.weight"], padding_idx=pad_id)
    x = x + sd[e + "token_type_embeddings.weight"][0]
    x = F.layer_norm(x, x.shape[-1:], sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], eps)
    B, L, E = x.shape
    dh = E // num_heads
    bias = torch.zeros(B, 1, 1, L, dtype=x.dtype, device=x.device)
    bias = bias.masked_fill(attention_mask[:, None, None, :] == 0, torch.finfo(x.dtype).min)
    for i in range(num_layers):
        p = f"{prefix}encoder.layer.{i}."
        q = F.linear(x, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"])
        k = F.linear(x, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"])
        v = F.linear(x, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"])
        q, k, v = (t.view(B, L, num_heads, dh).transpose(1, 2) for t in (q, k, v))
        att = (q @ k.transpose(-1, -2)) / math.sqrt(dh) + bias
        ctx = (att.softmax(-1) @ v).transpose(1, 2).reshape(B, L, E)
        y = F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
        x = F.layer_norm(y + x, (E,), sd[p + "attention.output.LayerNorm.weight"],
                         sd[p + "attention.output.LayerNorm.bias"], eps)
        h = F.gelu(F.linear(x, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
        y = F.linear(h, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
        x = F.layer_norm(y + x, (E,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], eps)
    return x


# ===================================================================================================== transformer


def mha(query: Tensor, key: Tensor, value: Tensor, sd: SD, prefix: str, nhead: int,
        key_padding_mask: Optional[Tensor]) -> Tensor:
    """torch.nn.MultiheadAttention forward (packed in_proj, q scaled by dh**-0.5, -inf key padding, softmax, PV,
    out_proj) as used at models/transformer.py:298,378,394.  Sequence-first tensors [S, B, E]."""
    Lq, B, E = query.shape
    Lk = key.shape[0]
    dh = E // nhead
    w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    q = q.reshape(Lq, B * nhead, dh).transpose(0, 1) * (dh ** -0.5)
    k = k.reshape(Lk, B * nhead, dh).transpose(0, 1)
    v = v.reshape(Lk, B * nhead, dh).transpose(0, 1)
    att = q @ k.transpose(1, 2)
    if key_padding_mask is not None:
        att = att.view(B, nhead, Lq, Lk).masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
        att = att.view(B * nhead, Lq, Lk)
    ctx = (att.softmax(-1) @ v).transpose(0, 1).reshape(Lq, B, E)
    return F.linear(ctx, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def _ln(x: Tensor, sd: SD, prefix: str, eps: float = 1e-5) -> Tensor:
    return F.layer_norm(x, x.shape[-1:], sd[prefix + "weight"], sd[prefix + "bias"], eps)


def _ffn(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.linear(F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])), sd[p + "linear2.weight"],
                    sd[p + "linear2.bias"])


def encoder(src: Tensor, mask: Tensor, pos: Tensor, sd: SD, prefix: str, num_layers: int, nhead: int) -> Tensor:
    """models/transformer.py:290-304 (post-norm layer) x num_layers; encoder.norm is None (transformer.py:46)."""
    x = src
    for i in range(num_layers):
        p = f"{prefix}layers.{i}."
        qk = x + pos
        x = _ln(x + mha(qk, qk, x, sd, p + "self_attn.", nhead, mask), sd, p + "norm1.")
        x = _ln(x + _ffn(x, sd, p), sd, p + "norm2.")
    return x


def decoder(tgt: Tensor, memory: Tensor, mask: Tensor, pos: Tensor, query_pos: Tensor, sd: SD, prefix: str,
            num_layers: int, nhead: int) -> Tensor:
    """models/transformer.py:362-408 x num_layers with the shared final norm applied to every layer's output
    (transformer.py:255-262).  Returns [num_layers, Q, B, E]."""
    x = tgt
    inter = []
    for i in range(num_layers):
        p = f"{prefix}layers.{i}."
        qk = x + query_pos
        x = _ln(x + mha(qk, qk, x, sd, p + "self_attn.", nhead, None), sd, p + "norm1.")
        x = _ln(x + mha(x + query_pos, memory + pos, memory, sd, p + "cross_attn_image.", nhead, mask), sd,
                p + "norm3.")
        x = _ln(x + _ffn(x, sd, p), sd, p + "norm4.")
        inter.append(_ln(x, sd, prefix + "norm."))
    return torch.stack(inter)


# ===================================================================================================== MDETR


class Config:
    """The hyper-parameters that define the arithmetic (main.py defaults)."""

    def __init__(self, backbone="resnet101", hidden_dim=256, nheads=8, enc_layers=6, dec_layers=6, num_queries=100,
                 masks=False, aux_loss=True, contrastive_align_loss=True, roberta_layers=12, roberta_heads=12,
                 roberta_eps=1e-12, set_cost_class=1.0, set_cost_bbox=5.0, set_cost_giou=2.0, eos_coef=0.1,
                 temperature_NCE=0.07, prefix=""):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def encode(sd: SD, cfg: Config, images: Tensor, pad_mask: Tensor, input_ids: Tensor, attention_mask: Tensor) -> dict:
    """Phase A: models/mdetr.py:377-394 (+ models/segmentation.py:59-80) and models/transformer.py:98-168."""
    P = cfg.prefix
    feats = resnet_body(images, sd, P + "backbone.0.body.", cfg.backbone)
    masks = [downsample_mask(pad_mask, f.shape[-2:]) for f in feats]
    src, mask = feats[-1], masks[-1]
    pos = position_sine(mask, cfg.hidden_dim // 2)
    src_proj = F.conv2d(src, sd[P + "input_proj.weight"], sd[P + "input_proj.bias"])
    bs = src.shape[0]
    src_seq = src_proj.flatten(2).permute(2, 0, 1)
    pos_seq = pos.flatten(2).permute(2, 0, 1)
    query_embed = sd[P + "query_embed.weight"].unsqueeze(1).repeat(1, bs, 1)
    text = roberta_encode(input_ids, attention_mask, sd, P + "transformer.text_encoder.", cfg.roberta_layers,
                          cfg.roberta_heads, cfg.roberta_eps).transpose(0, 1)
    text_attention_mask = attention_mask.ne(1).bool()
    r = P + "transformer.resizer."
    text_resized = F.layer_norm(F.linear(text, sd[r + "fc.weight"], sd[r + "fc.bias"]), (cfg.hidden_dim,),
                                sd[r + "layer_norm.weight"], sd[r + "layer_norm.bias"], 1e-12)
    src_all = torch.cat([src_seq, text_resized], 0)
    mask_all = torch.cat([mask.flatten(1), text_attention_mask], 1)
    pos_all = torch.cat([pos_seq, torch.zeros_like(text_resized)], 0)
    img_memory = encoder(src_all, mask_all, pos_all, sd, P + "transformer.encoder.", cfg.enc_layers, cfg.nheads)
    return {
        "text_memory_resized": text_resized,
        "text_memory": img_memory[-len(text_resized):],
        "img_memory": img_memory,
        "mask": mask_all,
        "text_attention_mask": text_attention_mask,
        "pos_embed": pos_all,
        "query_embed": query_embed,
        "features": feats,
        "feature_masks": masks,
        "src_proj": src_proj,
    }


def decode(sd: SD, cfg: Config, mc: dict, img_memory: Optional[Tensor] = None) -> dict:
    """Phase B: models/mdetr.py:396-462 (+ mask branch models/segmentation.py:157-167)."""
    P = cfg.prefix
    memory = mc["img_memory"] if img_memory is None else img_memory
    hs = decoder(torch.zeros_like(mc["query_embed"]), memory, mc["mask"], mc["pos_embed"], mc["query_embed"], sd,
                 P + "transformer.decoder.", cfg.dec_layers, cfg.nheads).transpose(1, 2)
    logits = F.linear(hs, sd[P + "class_embed.weight"], sd[P + "class_embed.bias"])
    x = hs
    for i in range(3):
        x = F.linear(x, sd[f"{P}bbox_embed.layers.{i}.weight"], sd[f"{P}bbox_embed.layers.{i}.bias"])
        if i < 2:
            x = F.relu(x)
    boxes = x.sigmoid()
    out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1], "hs": hs}
    if cfg.contrastive_align_loss:
        pq = F.normalize(F.linear(hs, sd[P + "contrastive_align_projection_image.weight"],
                                  sd[P + "contrastive_align_projection_image.bias"]), p=2, dim=-1)
        pt = F.normalize(F.linear(mc["text_memory"], sd[P + "contrastive_align_projection_text.weight"],
                                  sd[P + "contrastive_align_projection_text.bias"]).transpose(0, 1), p=2, dim=-1)
        out.update(proj_queries=pq[-1], proj_tokens=pt)
    if cfg.aux_loss:
        out["aux_outputs"] = []
        for l in range(cfg.dec_layers - 1):
            a = {"pred_logits": logits[l], "pred_boxes": boxes[l]}
            if cfg.contrastive_align_loss:
                a.update(proj_queries=pq[l], proj_tokens=pt)
            out["aux_outputs"].append(a)
    return out


# ===================================================================================================== mask branch


def attention_map(q: Tensor, k: Tensor, mask: Tensor, sd: SD, prefix: str, nheads: int) -> Tensor:
    """models/segmentation.py:262-273: per-head QK^T * dh**-0.5, -inf on padded pixels, softmax over H*W *per head*
    (flatten(3)), no value product.  q [B,Q,E], k [B,E,H,W] -> [B,Q,heads,H,W]."""
    E = q.shape[-1]
    q = F.linear(q, sd[prefix + "q_linear.weight"], sd[prefix + "q_linear.bias"])
    k = F.conv2d(k, sd[prefix + "k_linear.weight"][:, :, None, None], sd[prefix + "k_linear.bias"])
    dh = E // nheads
    qh = q.view(q.shape[0], q.shape[1], nheads, dh)
    kh = k.view(k.shape[0], nheads, dh, k.shape[-2], k.shape[-1])
    w = torch.einsum("bqnc,bnchw->bqnhw", qh * (float(dh) ** -0.5), kh)
    w = w.masked_fill(mask[:, None, None], float("-inf"))
    return F.softmax(w.flatten(3), dim=-1).view_as(w)


def mask_head(src_proj: Tensor, bbox_mask: Tensor, fpns: Sequence[Tensor], sd: SD, prefix: str) -> Tensor:
    """models/segmentation.py:203-241 (MaskHeadSmallConv): 5 x (3x3 conv, GroupNorm(8), ReLU) with FPN adapters and
    nearest upsampling, then a 3x3 conv to one channel.  Returns [B*Q, 1, H1, W1]."""
    nq = bbox_mask.shape[1]

    def expand(t, n):
        return t.unsqueeze(1).repeat(1, int(n), 1, 1, 1).flatten(0, 1)

    def block(x, i):
        x = F.conv2d(x, sd[f"{prefix}lay{i}.weight"], sd[f"{prefix}lay{i}.bias"], padding=1)
        return F.relu(F.group_norm(x, 8, sd[f"{prefix}gn{i}.weight"], sd[f"{prefix}gn{i}.bias"]))

    x = torch.cat([expand(src_proj, nq), bbox_mask.flatten(0, 1)], 1)
    x = block(block(x, 1), 2)
    for i, fpn in enumerate(fpns, start=1):
        cur = F.conv2d(fpn, sd[f"{prefix}adapter{i}.weight"], sd[f"{prefix}adapter{i}.bias"])
        cur = expand(cur, x.shape[0] // cur.shape[0])
        x = cur + F.interpolate(x, size=cur.shape[-2:], mode="nearest")
        x = block(x, i + 2)
    return F.conv2d(x, sd[prefix + "out_lay.weight"], sd[prefix + "out_lay.bias"], padding=1)


def decode_masks(sd: SD, cfg: Config, mc: dict, out: dict, seg_prefix: str = "") -> Tensor:
    """models/segmentation.py:157-167: pred_masks [B, Q, H1, W1]."""
    feats, fmasks = mc["features"], mc["feature_masks"]
    src_proj = mc["src_proj"]
    bs = src_proj.shape[0]
    n_text = len(mc["text_memory"])
    memory = mc["img_memory"][:-n_text].permute(1, 2, 0).view_as(src_proj)
    bbox_mask = attention_map(out["hs"][-1], memory, fmasks[-1], sd, seg_prefix + "bbox_attention.", cfg.nheads)
    seg = mask_head(src_proj, bbox_mask, [feats[2], feats[1], feats[0]], sd, seg_prefix + "mask_head.")
    return seg.view(bs, cfg.num_queries, seg.shape[-2], seg.shape[-1])


# ===================================================================================================== criterion


def _src_idx(indices):
    b = torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(indices)])
    return b, torch.cat([s for s, _ in indices])


def loss_labels(logits: Tensor, targets, positive_map: Tensor, indices, num_boxes: float, eos_coef: float) -> Tensor:
    """models/mdetr.py:488-518 soft-token cross entropy."""
    logp = logits.log_softmax(-1)
    src = _src_idx(indices)
    tgt_idx, off = [], 0
    for i, (_, t) in enumerate(indices):
        tgt_idx.append(t + off)
        off += len(targets[i]["boxes"])
    tgt_idx = torch.cat(tgt_idx)
    sim = torch.zeros_like(logp)
    sim[:, :, -1] = 1
    sim[src] = positive_map[tgt_idx]
    ce = -(logp * sim).sum(-1)
    wgt = torch.full(ce.shape, eos_coef, device=ce.device)
    wgt[src] = 1
    return (ce * wgt).sum() / num_boxes


def loss_boxes(pred_boxes: Tensor, targets, indices, num_boxes: float) -> Tuple[Tensor, Tensor]:
    """models/mdetr.py:805-825: (L1, 1 - diag GIoU), each summed / num_boxes."""
    src = pred_boxes[_src_idx(indices)]
    tgt = torch.cat([t["boxes"][i] for t, (_, i) in zip(targets, indices)], 0)
    l1 = F.l1_loss(src, tgt, reduction="none").sum() / num_boxes
    giou = (1 - torch.diag(pairwise_giou(box_cxcywh_to_xyxy(src), box_cxcywh_to_xyxy(tgt)))).sum() / num_boxes
    return l1, giou


def loss_cardinality(logits: Tensor, targets) -> Tensor:
    """models/mdetr.py:783-803 (logging only)."""
    lengths = torch.as_tensor([len(t["labels"]) for t in targets], device=logits.device)
    card = (logits.argmax(-1) != logits.shape[-1] - 1).sum(1)
    return F.l1_loss(card.float(), lengths.float())


def token_spans(tokenized, i: int, spans) -> List[Tuple[int, int]]:
    """The char-span -> token-span resolution shared by models/mdetr.py:622-643 (and :119-141, :183-205)."""
    res = []
    for beg, end in spans:
        b = tokenized.char_to_token(i, beg)
        e = tokenized.char_to_token(i, end - 1)
        if b is None:
            try:
                b = tokenized.char_to_token(beg + 1)
                if b is None:
                    b = tokenized.char_to_token(beg + 2)
            except Exception:
                b = None
        if e is None:
            try:
                e = tokenized.char_to_token(end - 2)
                if e is None:
                    e = tokenized.char_to_token(end - 3)
            except Exception:
                e = None
        if b is None or e is None:
            continue
        res.append((b, e))
    return res


def loss_contrastive_align(proj_queries: Tensor, proj_tokens: Tensor, tokenized, targets, indices, num_boxes: float,
                           temperature: float) -> Tensor:
    """models/mdetr.py:601-666."""
    logits = proj_queries @ proj_tokens.transpose(-1, -2) / temperature
    pm = torch.zeros(logits.shape, dtype=torch.bool)
    for i, ((isrc, itgt), tgt) in enumerate(zip(indices, targets)):
        key = "tokens_positive" if "tokens_positive" in tgt else "tokens"
        for j, t in enumerate(itgt.tolist()):
            for b, e in token_spans(tokenized, i, tgt[key][t]):
                pm[i, isrc[j], b:e + 1] = True
    pm = pm.to(logits.device)
    pos = -logits.masked_fill(~pm, 0)
    b2t = ((pos.sum(2) / (pm.sum(2) + 1e-6) + logits.logsumexp(2))).masked_fill(~pm.any(2), 0).sum()
    t2b = ((pos.sum(1) / (pm.sum(1) + 1e-6) + logits.logsumexp(1))).masked_fill(~pm.any(1), 0).sum()
    return (b2t + t2b) / 2 / num_boxes


def sigmoid_focal_loss(x: Tensor, t: Tensor, num_boxes: float, alpha: float = 0.25, gamma: float = 2) -> Tensor:
    """models/segmentation.py:294-319."""
    p = x.sigmoid()
    ce = F.binary_cross_entropy_with_logits(x, t, reduction="none")
    p_t = p * t + (1 - p) * (1 - t)
    loss = ce * ((1 - p_t) ** gamma)
    loss = (alpha * t + (1 - alpha) * (1 - t)) * loss
    return loss.mean(1).sum() / num_boxes


def dice_loss(x: Tensor, t: Tensor, num_boxes: float) -> Tensor:
    """models/segmentation.py:276-291."""
    p = x.sigmoid().flatten(1)
    num = 2 * (p * t).sum(1)
    den = p.sum(-1) + t.sum(-1)
    return (1 - (num + 1) / (den + 1)).sum() / num_boxes


def loss_masks(pred_masks: Tensor, targets, indices, num_boxes: float) -> Tuple[Tensor, Tensor]:
    """models/mdetr.py:827-853: bilinear upsample of matched masks to the padded target size, focal + dice."""
    src = _src_idx(indices)
    tb = torch.cat([torch.full_like(t, i) for i, (_, t) in enumerate(indices)])
    ti = torch.cat([t for _, t in indices])
    tms = [t["masks"] for t in targets]
    H = max(m.shape[-2] for m in tms)
    W = max(m.shape[-1] for m in tms)
    nmax = max(m.shape[0] for m in tms)
    tgt = torch.zeros(len(tms), nmax, H, W, dtype=pred_masks.dtype, device=pred_masks.device)
    for i, m in enumerate(tms):
        tgt[i, : m.shape[0], : m.shape[1], : m.shape[2]] = m.to(pred_masks)
    sm = pred_masks[src]
    sm = F.interpolate(sm[:, None], size=(H, W), mode="bilinear", align_corners=False)[:, 0].flatten(1)
    tm = tgt[tb, ti].flatten(1)
    return sigmoid_focal_loss(sm, tm, num_boxes), dice_loss(sm, tm, num_boxes)


def criterion(cfg: Config, out: dict, tokenized, targets, positive_map: Tensor, world_size: int = 1,
              num_boxes: Optional[float] = None, masks: bool = False,
              forced_indices: Optional[list] = None) -> Tuple[Dict[str, Tensor], list]:
    """models/mdetr.py:990-1021 (single-model branch). Returns (losses, indices of every decoder layer), ordered
    [main, aux_0, aux_1, ...].  `forced_indices` (same order) replaces the matcher: used by gradient-parity tests so
    that both implementations differentiate the same assignment when near-tie costs would otherwise flip it."""
    def one(o, suffix, with_masks, forced=None):
        idx = forced if forced is not None else hungarian_match(
            o["pred_logits"], o["pred_boxes"], targets, positive_map, cfg.set_cost_class, cfg.set_cost_bbox,
            cfg.set_cost_giou)
        res = {"loss_ce" + suffix: loss_labels(o["pred_logits"], targets, positive_map, idx, nb, cfg.eos_coef)}
        l1, gi = loss_boxes(o["pred_boxes"], targets, idx, nb)
        res["loss_bbox" + suffix], res["loss_giou" + suffix] = l1, gi
        res["cardinality_error" + suffix] = loss_cardinality(o["pred_logits"], targets)
        if with_masks:
            res["loss_mask" + suffix], res["loss_dice" + suffix] = loss_masks(o["pred_masks"], targets, idx, nb)
        if cfg.contrastive_align_loss:
            # differentiable in the reference: models/mdetr.py:601 is a plain method (only softkd_matcher :520 and
            # loss_cardinality :783 carry @torch.no_grad()); weight 1 per layer (:1068-1069), summed at engine.py:72
            res["loss_contrastive_align" + suffix] = loss_contrastive_align(
                o["proj_queries"], o["proj_tokens"], tokenized, targets, idx, nb, cfg.temperature_NCE)
        return res, idx

    nb = num_boxes
    if nb is None:
        nb = max(float(sum(len(t["labels"]) for t in targets)) / world_size, 1.0)
    losses, idx = one(out, "", masks, forced_indices[0] if forced_indices else None)
    all_idx = [idx]
    for i, aux in enumerate(out.get("aux_outputs", [])):
        l, idx_i = one(aux, f"_{i}", False, forced_indices[i + 1] if forced_indices else None)
        losses.update(l)
        all_idx.append(idx_i)
    return losses, all_idx


# ===================================================================================================== distillation


def softkd_match(prob_t: Tensor, prob_s: Tensor, box_t: Tensor, box_s: Tensor) -> Tuple[Tensor, Tensor]:
    """models/mdetr.py:520-541: unweighted KL + L1 - GIoU cost, rows = student (source), cols = teacher."""
    with torch.no_grad():
        cost_class = (prob_t * (prob_t.unsqueeze(0).log() - prob_s.log().unsqueeze(1))).sum(-1)
        cost_bbox = torch.cdist(box_s, box_t, p=1)
        cost_giou = -pairwise_giou(box_cxcywh_to_xyxy(box_s), box_cxcywh_to_xyxy(box_t))
        r, c = lsap((cost_bbox + cost_class + cost_giou).cpu().numpy())
        return torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(c, dtype=torch.int64)


def loss_softkd(out_noun: dict, out_sth: dict, idx_noun, idx_sth, num_queries: int) -> Tensor:
    """models/mdetr.py:543-599."""
    pn = out_noun["pred_logits"].detach().softmax(-1)
    ps = out_sth["pred_logits"].softmax(-1)
    bn = torch.cat([pn[..., :-1].sum(-1, keepdim=True), pn[..., -1:]], -1)
    bs_ = torch.cat([ps[..., :-1].sum(-1, keepdim=True), ps[..., -1:]], -1)
    total = torch.tensor(0.0, device=pn.device)
    for i in range(len(idx_noun)):
        tn, ts = idx_noun[i], idx_sth[i]
        tp_n = torch.zeros(tn[0].shape[0], 2, device=pn.device)
        tp_s = torch.zeros(ts[0].shape[0], 2, device=pn.device)
        tp_n[tn[1]] = bn[i][tn[0]]
        tp_s[ts[1]] = bs_[i][ts[0]]
        fn = torch.ones(num_queries, dtype=torch.bool, device=pn.device)
        fn[tn[0]] = False
        fs = torch.ones(num_queries, dtype=torch.bool, device=pn.device)
        fs[ts[0]] = False
        fp_n, fp_s = bn[i][fn], bs_[i][fs]
        r, c = softkd_match(fp_n, fp_s, out_noun["pred_boxes"][i][fn], out_sth["pred_boxes"][i][fs])
        ln = torch.cat([tp_n, fp_n[c]], 0)
        ls = torch.cat([tp_s, fp_s[r]], 0)
        total = total + F.kl_div(ls.log(), ln, reduction="batchmean")
    return total / len(idx_noun)


# ===================================================================================================== k-means


def kmeans(X: Tensor, init_centers: Tensor, num_clusters: int, full_label: float, tol: float = 1e-4,
           rng: Optional[np.random.RandomState] = None) -> Tuple[Tensor, Tensor]:
    """models/kmeans.py:21-96: Lloyd iterations until (sum of centre shifts)**2 < tol; empty clusters keep their
    centre; random initial rows (numpy RNG, kmeans.py:16) unless `full_label`."""
    X = X.float()
    if full_label == 0:
        pick = (rng or np.random).choice(len(X), num_clusters, replace=False)
        centers = X[pick]
    else:
        centers = init_centers
    while True:
        dis = ((X.unsqueeze(1) - centers.unsqueeze(0)) ** 2.0).sum(-1)
        choice = torch.argmin(dis, dim=1)
        prev = centers.clone()
        for k in range(num_clusters):
            sel = X[choice == k]
            if len(sel) != 0:
                centers[k] = sel.mean(0)
        shift = torch.sum(torch.sqrt(torch.sum((centers - prev) ** 2, dim=1)))
        if shift ** 2 < tol:
            break
    return choice, centers


def kmeans_predict(X: Tensor, centers: Tensor) -> Tensor:
    """models/kmeans.py:99-133."""
    return torch.argmin(((X.float().unsqueeze(1) - centers.unsqueeze(0)) ** 2.0).sum(-1), dim=1)
