"""Forward / backward of the model's blocks written directly against the sm_100a kernels (no autograd in here).

Every `*_fwd` returns `(output, saved)`; the matching `*_bwd` consumes `saved` and the output gradient, returns the
input gradient(s) and stores parameter gradients (fp32, in the parameter's own layout) into the dict `g` under the
parameter's name.  `w` maps a parameter name to the tensor the kernels read: the bf16 *shadow* for matrices (kept
fresh by kernels.WeightPrep, FrozenBatchNorm scales folded in) and the fp32 parameter itself for vectors.  `req`
is the set of names whose gradient is wanted.

Activations are bf16: [rows, features] for sequences (row = s * B + b, i.e. the reference's [S, B, C] layout) and NHWC
for the convolutional trunk.  Accumulation, LayerNorm statistics, softmax and every loss are fp32.

Reference semantics: models/transformer.py:270-470 (encoder / decoder layers), transformers' RobertaLayer,
torchvision Bottleneck + models/backbone.py:21-58, models/mdetr.py:420-433 (heads).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Set, Tuple

import torch

from . import kernels as K
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, GEMM_DGRAD, GEMM_FWD, GEMM_WGRAD

BF = torch.bfloat16
W = Dict[str, torch.Tensor]
G = Dict[str, torch.Tensor]


class zero_arena:
    """`with zero_arena(numel, device):` serves every parameter-gradient buffer of a backward stage from ONE zeroed
    fp32 allocation (one memset instead of several hundred); falls back to torch.zeros when exhausted."""

    current = None
    mark_bytes = int(os.environ.get("TOIST_GRAD_MARK_MB", "64")) << 20

    def __init__(self, numel: int, dev, marks: bool = False):
        self.buf = torch.zeros(max(int(numel), 1), dtype=torch.float32, device=dev)
        self.off = 0
        self.want_marks = marks
        self.marks = []   # (offset in elements, event): everything below the offset is final once the event has fired
        self.last = 0

    def __enter__(self):
        self.prev, zero_arena.current = zero_arena.current, self
        return self

    def __exit__(self, *exc):
        zero_arena.current = self.prev
        return False

    def take(self, shape, dev):
        n = 1
        for d in shape:
            n *= int(d)
        end = self.off + n
        if end > self.buf.numel() or self.buf.device != torch.device(dev):
            return torch.zeros(shape, dtype=torch.float32, device=dev)
        v = self.buf[self.off:end].view(shape)
        self.off = (end + 63) // 64 * 64  # keep every buffer 256-byte aligned (vector epilogues, TMA)
        return v


def mark_grads() -> None:
    """Called by the backward loops between layers / blocks.  When the data-parallel exchange asked for it, records an
    event once another `mark_bytes` of parameter gradients are complete: gradients are served from the arena in the
    order the backward produces them, so "complete" is a prefix of the arena and util/dist.FlatGradSync can all-reduce
    that prefix while the rest of the stage is still running.  Weight-gradient kernels run on the side lane
    (kernels.wgrad_lane); the event is recorded there, after the lane has been made to wait for the main stream, so it
    covers both.  Inside a CUDA-graph capture the event is an external event-record node."""
    a = zero_arena.current
    if a is None or not a.want_marks or (a.off - a.last) * 4 < zero_arena.mark_bytes:
        return
    main = torch.cuda.current_stream()
    lane = K._WgradLane
    stream = main
    if lane.active and lane.dirty and lane.stream is not None:
        lane.stream.wait_stream(main)
        for s_ in lane.also_wait:
            lane.stream.wait_stream(s_)
        stream = lane.stream
    else:
        for s_ in lane.also_wait:
            main.wait_stream(s_)
    ev = torch.cuda.Event(external=torch.cuda.is_current_stream_capturing())
    ev.record(stream)
    a.marks.append((a.off, ev))
    a.last = a.off


def _zeros(shape, dev):
    shape = tuple(shape) if isinstance(shape, (tuple, list, torch.Size)) else (shape,)
    a = zero_arena.current
    if a is not None:
        return a.take(shape, dev)
    return torch.zeros(shape, dtype=torch.float32, device=dev)


class Drop:
    """Dropout context of one block: probability, device seed tensor and the block's first site id.  `site(k)` is the
    (p, seed, site) triple of the k-th dropout application inside the block; forward and backward use the same k."""

    __slots__ = ("p", "seed", "base")

    def __init__(self, p: float, seed: torch.Tensor, base: int):
        self.p, self.seed, self.base = float(p), seed, int(base)

    def site(self, k: int):
        return (self.p, self.seed, self.base + k)

    @property
    def keep_scale(self) -> float:
        return 1.0 / (1.0 - self.p)


def _site(drop: Optional["Drop"], k: int):
    return None if drop is None else drop.site(k)


def lin_drop_res(x, wt, bias, res, drop_site):
    """res + dropout(linear(x)): the residual branches of the transformer layers (dropout1/2/3/4, HF hidden dropout)."""
    if drop_site is None:
        return K.linear_fwd(x, wt, bias, res=res)
    y = K.linear_fwd(x, wt, bias)
    return K.dropout(y, drop_site, res=res, out=y)


def drop_grad(ds, drop_site):
    """Gradient w.r.t. the linear output behind `lin_drop_res` (ds = gradient of the sum)."""
    return ds if drop_site is None else K.dropout(ds, drop_site)


# ------------------------------------------------------------------------------------------------ linear pieces
def lin_param_grads(g: G, req: Set[str], wname: str, bname: Optional[str], dy: torch.Tensor, x: torch.Tensor,
                    w_shape) -> None:
    """dW = dy^T x, db = column sums of dy (nn.Linear backward)."""
    with K.wgrad_lane(dy, x):
        if wname in req:
            dw = _zeros(tuple(w_shape), dy.device)
            K.linear_wgrad(dy, x, dw)
            g[wname] = dw
        if bname is not None and bname in req:
            db = _zeros((dy.shape[1],), dy.device)
            K.colsum(dy, db)
            g[bname] = db


def ln_fwd(w: W, pre: str, x: torch.Tensor, eps: float, **kw):
    return K.layernorm_fwd(x, w[pre + "weight"], w[pre + "bias"], eps, **kw)


def ln_bwd(w: W, g: G, req: Set[str], pre: str, dy, x, mean, rstd, dy2=None, dx_dtype=BF):
    """Accumulates into g[pre+weight/bias] (several LayerNorm applications may share parameters)."""
    dg = db = None
    if (pre + "weight") in req:
        if (pre + "weight") not in g:
            g[pre + "weight"] = _zeros((x.shape[1],), x.device)
            g[pre + "bias"] = _zeros((x.shape[1],), x.device)
        dg, db = g[pre + "weight"], g[pre + "bias"]
    return K.layernorm_bwd(dy, x, mean, rstd, w[pre + "weight"], dy2=dy2, dgamma=dg, dbeta=db, dx_dtype=dx_dtype)


def branch_ln_fwd(w: W, x_in, wname: str, bname: str, res, drop_site, ln_pre: str, eps: float, add=None):
    """One residual branch: LayerNorm(res + dropout(linear(x_in))) and, with `add`, the same plus `add` (the positional
    / query embedding the next attention adds to its input).  Returns (out, out_add | None, s, mean, rstd) where s is
    the LayerNorm input the backward needs.  Training mode (dropout) and / or `add` go through the fused kernel
    (csrc/norm.cu: one launch instead of dropout + LayerNorm + add); results are bit-identical to the separate kernels."""
    gamma, beta = w[ln_pre + "weight"], w[ln_pre + "bias"]
    if drop_site is None:
        s = K.linear_fwd(x_in, w[wname], w[bname], res=res)  # the residual add is the GEMM's epilogue
        if add is not None and K.fused_ln_ok(s, add):
            out, out_add, _, m, r = K.layernorm_fused_fwd(s, None, add, gamma, beta, eps, None, want_sum=False)
            return out, out_add, s, m, r
        out, _, m, r = K.layernorm_fwd(s, gamma, beta, eps)
        return out, (K.add_bf16(out, add) if add is not None else None), s, m, r
    y = K.linear_fwd(x_in, w[wname], w[bname])
    if K.fused_ln_ok(y, res, add):
        return K.layernorm_fused_fwd(y, res, add, gamma, beta, eps, drop_site)
    s = K.dropout(y, drop_site, res=res, out=y)
    out, _, m, r = K.layernorm_fwd(s, gamma, beta, eps)
    return out, (K.add_bf16(out, add) if add is not None else None), s, m, r


def branch_ln_bwd(w: W, g: G, req: Set[str], ln_pre: str, dy, s, mean, rstd, drop_site, dy2=None):
    """Backward of `branch_ln_fwd` up to the linear layer's output: returns (ds, d_linear_out) with ds the gradient of
    the residual sum (it also flows into the residual input) and d_linear_out = dropout-backward(ds)."""
    if drop_site is None:
        ds = ln_bwd(w, g, req, ln_pre, dy, s, mean, rstd, dy2=dy2)
        return ds, ds
    if K.fused_ln_ok(dy, dy2, s):
        dg = db = None
        if (ln_pre + "weight") in req:
            if (ln_pre + "weight") not in g:
                g[ln_pre + "weight"] = _zeros((s.shape[1],), s.device)
                g[ln_pre + "bias"] = _zeros((s.shape[1],), s.device)
            dg, db = g[ln_pre + "weight"], g[ln_pre + "bias"]
        return K.layernorm_bwd_drop(dy, s, mean, rstd, w[ln_pre + "weight"], drop_site, dy2=dy2, dgamma=dg, dbeta=db)
    ds = ln_bwd(w, g, req, ln_pre, dy, s, mean, rstd, dy2=dy2)
    return ds, K.dropout(ds, drop_site)


# ------------------------------------------------------------------------------------------------ attention
def mha_fwd(w: W, pre: str, xq, xk, xv, key_mask, nhead: int, B: int, drop_site=None, kv=None):
    """nn.MultiheadAttention up to (excluding) out_proj.  xq/xk/xv: [rows, E] bf16; xq is xk -> fused q|k GEMM.
    `kv` = (k2, v2): key / value projections computed by the caller (the decoder projects the encoder memory for all of
    its layers in one GEMM each, runtime.decoder_fwd); column slices of wider tensors are fine."""
    E = xq.shape[1]
    Wi, bi = w[pre + "in_proj_weight"], w[pre + "in_proj_bias"]
    if kv is not None:
        q2 = K.linear_fwd(xq, Wi[:E], bi[:E])
        k2, v2 = kv
        Sq, Sk = xq.shape[0] // B, xk.shape[0] // B
        ctx, probs = K.attention_fwd(q2.view(Sq, B, E), k2.view(Sk, B, E), v2.view(Sk, B, E), key_mask, nhead,
                                     drop=drop_site)
        return ctx.view(Sq * B, E), (xq, xk, xv, q2, k2, v2, probs)
    if xq is xk:
        qk = K.linear_fwd(xq, Wi[: 2 * E], bi[: 2 * E])
        q2, k2 = qk[:, :E], qk[:, E:]
    else:
        q2 = K.linear_fwd(xq, Wi[:E], bi[:E])
        k2 = K.linear_fwd(xk, Wi[E: 2 * E], bi[E: 2 * E])
    v2 = K.linear_fwd(xv, Wi[2 * E:], bi[2 * E:])
    Sq, Sk = xq.shape[0] // B, xk.shape[0] // B
    ctx, probs = K.attention_fwd(q2.view(Sq, B, E), k2.view(Sk, B, E), v2.view(Sk, B, E), key_mask, nhead,
                                 drop=drop_site)
    return ctx.view(Sq * B, E), (xq, xk, xv, q2, k2, v2, probs)


def mha_bwd(w: W, g: G, req: Set[str], pre: str, dctx, saved, nhead: int, B: int, need=(True, True, True),
            drop_site=None, dkv=None):
    """Returns (dxq, dxk, dxv); for self-attention (xq is xk) dxq is the gradient of the shared input and dxk None.
    `dkv` = (dk2, dv2): buffers (column slices) that receive the gradients of the key / value projections when the
    caller projected them itself (see mha_fwd); their data gradients are then the caller's business (need[1:] False)."""
    xq, xk, xv, q2, k2, v2, probs = saved
    E = xq.shape[1]
    Sq, Sk = xq.shape[0] // B, xk.shape[0] // B
    dev = xq.device
    Wi = w[pre + "in_proj_weight"]
    fused = xq is xk
    if fused:
        dqk = torch.empty((xq.shape[0], 2 * E), dtype=BF, device=dev)
        dq2, dk2 = dqk[:, :E], dqk[:, E:]
    else:
        dq2 = torch.empty((xq.shape[0], E), dtype=BF, device=dev)
        dk2 = dkv[0] if dkv is not None else torch.empty((xk.shape[0], E), dtype=BF, device=dev)
    dv2 = dkv[1] if dkv is not None else torch.empty((xv.shape[0], E), dtype=BF, device=dev)
    K.attention_bwd(dctx.view(Sq, B, E), q2.view(Sq, B, E), k2.view(Sk, B, E), v2.view(Sk, B, E), probs, nhead,
                    dq2.view(Sq, B, E), dk2.view(Sk, B, E), dv2.view(Sk, B, E), drop=drop_site)
    wn, bn = pre + "in_proj_weight", pre + "in_proj_bias"
    with K.wgrad_lane(dq2, dk2, dv2, xq, xk, xv, dqk if fused else None):
        if wn in req:
            dw = _zeros((3 * E, E), dev)
            if fused:
                K.linear_wgrad(dqk, xq, dw[: 2 * E])
            else:
                K.linear_wgrad(dq2, xq, dw[:E])
                K.linear_wgrad(dk2, xk, dw[E: 2 * E])
            K.linear_wgrad(dv2, xv, dw[2 * E:])
            g[wn] = dw
        if bn in req:
            db = _zeros((3 * E,), dev)
            if fused:
                K.colsum(dqk, db[: 2 * E])
            else:
                K.colsum(dq2, db[:E])
                K.colsum(dk2, db[E: 2 * E])
            K.colsum(dv2, db[2 * E:])
            g[bn] = db
    dxq = dxk = dxv = None
    if fused:
        if need[0] or need[1]:
            dxq = K.linear_dgrad(dqk, Wi[: 2 * E])
    else:
        if need[0]:
            dxq = K.linear_dgrad(dq2, Wi[:E])
        if need[1]:
            dxk = K.linear_dgrad(dk2, Wi[E: 2 * E])
    if need[2]:
        dxv = K.linear_dgrad(dv2, Wi[2 * E:])
    return dxq, dxk, dxv


def _ffn_hidden(w: W, x, drop: Optional[Drop], k_hidden: int):
    """dropout(relu(linear1(x))).  The hidden dropout is applied in place, so the saved `h` is zero exactly where
    either the ReLU or the dropout mask is zero."""
    h = K.linear_fwd(x, w["linear1.weight"], w["linear1.bias"], act=ACT_RELU)
    if drop is not None:
        K.dropout(h, drop.site(k_hidden), out=h)
    return h


def _ffn_bwd(w: W, g: G, req: Set[str], ds, dy, x, h, drop: Optional[Drop] = None):
    """ds: gradient of the FFN sum (residual path), dy: gradient of linear2's output (= dropout-backward of ds);
    returns dx including the residual path."""
    lin_param_grads(g, req, "linear2.weight", "linear2.bias", dy, h, w["linear2.weight"].shape)
    # (h > 0) is the product of the ReLU and hidden-dropout masks; the surviving elements carry the 1/(1-p) factor
    dh = K.linear_dgrad(dy, w["linear2.weight"], mask=h, alpha=1.0 if drop is None else drop.keep_scale)
    lin_param_grads(g, req, "linear1.weight", "linear1.bias", dh, x, w["linear1.weight"].shape)
    return K.linear_dgrad(dh, w["linear1.weight"], res=ds)


# ------------------------------------------------------------------------------------------------ encoder layer
def encoder_layer_fwd(w: W, x, pos, key_mask, nhead: int, B: int, drop: Optional[Drop] = None, xp=None,
                      want_next_xp: bool = False):
    """models/transformer.py:290-304 (post-norm).  x, pos [S*B, E] bf16.  Dropout sites (training): 0 attention
    weights, 1 dropout1, 2 FFN hidden, 3 dropout2.  `xp` = x + pos when the previous layer already produced it;
    with `want_next_xp` the output + pos for the next layer comes out of the last LayerNorm launch."""
    if xp is None:
        xp = K.add_bf16(x, pos)
    ctx, sv = mha_fwd(w, "self_attn.", xp, xp, x, key_mask, nhead, B, _site(drop, 0))
    x1, _, s1, m1, r1 = branch_ln_fwd(w, ctx, "self_attn.out_proj.weight", "self_attn.out_proj.bias", x, _site(drop, 1),
                                      "norm1.", 1e-5)
    h = _ffn_hidden(w, x1, drop, 2)
    x2, xp_next, s2, m2, r2 = branch_ln_fwd(w, h, "linear2.weight", "linear2.bias", x1, _site(drop, 3), "norm2.", 1e-5,
                                            add=pos if want_next_xp else None)
    return x2, (sv, ctx, s1, m1, r1, x1, h, s2, m2, r2), xp_next


def encoder_layer_bwd(w: W, g: G, req: Set[str], dy, saved, nhead: int, B: int, drop: Optional[Drop] = None, dy2=None):
    """dy (+ dy2): gradient of the layer output (dy2: the part that arrived through the next layer's x + pos input)."""
    sv, ctx, s1, m1, r1, x1, h, s2, m2, r2 = saved
    ds2, dlin2 = branch_ln_bwd(w, g, req, "norm2.", dy, s2, m2, r2, _site(drop, 3), dy2=dy2)
    dx1 = _ffn_bwd(w, g, req, ds2, dlin2, x1, h, drop)
    ds1, dy1 = branch_ln_bwd(w, g, req, "norm1.", dx1, s1, m1, r1, _site(drop, 1))
    lin_param_grads(g, req, "self_attn.out_proj.weight", "self_attn.out_proj.bias", dy1, ctx,
                    w["self_attn.out_proj.weight"].shape)
    dctx = K.linear_dgrad(dy1, w["self_attn.out_proj.weight"])
    dxp, _, dxv = mha_bwd(w, g, req, "self_attn.", dctx, sv, nhead, B, drop_site=_site(drop, 0))
    # residual + v path = gradient of x itself; dxp = gradient of (x + pos), handed over separately so that the caller
    # can feed it to the previous layer's LayerNorm backward as its second gradient (pos carries none: sine embedding)
    return K.add_bf16(ds1, dxv), dxp


# ------------------------------------------------------------------------------------------------ decoder layer
def decoder_layer_fwd(w: W, tgt, qpos, mem, mem_pos, key_mask, nhead: int, B: int, drop: Optional[Drop] = None,
                      tq=None, want_next_tq: bool = False, kv=None):
    """models/transformer.py:362-408 (post-norm; the text cross-attention is disabled in the reference).  Dropout
    sites: 0 self-attention weights, 1 dropout1, 2 cross-attention weights, 3 dropout3, 4 FFN hidden, 5 dropout4.
    `tq` = tgt + qpos when the previous layer already produced it (want_next_tq)."""
    if tq is None:
        tq = K.add_bf16(tgt, qpos)
    ctx1, sv1 = mha_fwd(w, "self_attn.", tq, tq, tgt, None, nhead, B, _site(drop, 0))
    t1, cq, s1, m1, r1 = branch_ln_fwd(w, ctx1, "self_attn.out_proj.weight", "self_attn.out_proj.bias", tgt,
                                       _site(drop, 1), "norm1.", 1e-5, add=qpos)
    ctx2, sv2 = mha_fwd(w, "cross_attn_image.", cq, mem_pos, mem, key_mask, nhead, B, _site(drop, 2), kv=kv)
    t2, _, s2, m2, r2 = branch_ln_fwd(w, ctx2, "cross_attn_image.out_proj.weight", "cross_attn_image.out_proj.bias", t1,
                                      _site(drop, 3), "norm3.", 1e-5)
    h = _ffn_hidden(w, t2, drop, 4)
    t3, tq_next, s3, m3, r3 = branch_ln_fwd(w, h, "linear2.weight", "linear2.bias", t2, _site(drop, 5), "norm4.", 1e-5,
                                            add=qpos if want_next_tq else None)
    return t3, (sv1, ctx1, s1, m1, r1, t1, sv2, ctx2, s2, m2, r2, t2, h, s3, m3, r3), tq_next


def decoder_layer_bwd(w: W, g: G, req: Set[str], dy, dy2, saved, nhead: int, B: int, need_tgt: bool = True,
                      drop: Optional[Drop] = None, dkv=None):
    """dy (+ dy2): gradient of the layer output.  Returns (d_tgt, d_qpos, d_mem_pos, d_mem); with `dkv` (hoisted key /
    value projections of the memory, see mha_fwd) the last two are None and dkv's buffers hold dK / dV."""
    sv1, ctx1, s1, m1, r1, t1, sv2, ctx2, s2, m2, r2, t2, h, s3, m3, r3 = saved
    ds3, dlin3 = branch_ln_bwd(w, g, req, "norm4.", dy, s3, m3, r3, _site(drop, 5), dy2=dy2)
    dt2 = _ffn_bwd(w, g, req, ds3, dlin3, t2, h, drop)
    ds2, dy2_ = branch_ln_bwd(w, g, req, "norm3.", dt2, s2, m2, r2, _site(drop, 3))
    lin_param_grads(g, req, "cross_attn_image.out_proj.weight", "cross_attn_image.out_proj.bias", dy2_, ctx2,
                    w["cross_attn_image.out_proj.weight"].shape)
    dctx2 = K.linear_dgrad(dy2_, w["cross_attn_image.out_proj.weight"])
    dcq, dmem_pos, dmem = mha_bwd(w, g, req, "cross_attn_image.", dctx2, sv2, nhead, B, drop_site=_site(drop, 2),
                                  need=(True, dkv is None, dkv is None), dkv=dkv)
    ds1, dy1 = branch_ln_bwd(w, g, req, "norm1.", ds2, s1, m1, r1, _site(drop, 1), dy2=dcq)
    lin_param_grads(g, req, "self_attn.out_proj.weight", "self_attn.out_proj.bias", dy1, ctx1,
                    w["self_attn.out_proj.weight"].shape)
    dctx1 = K.linear_dgrad(dy1, w["self_attn.out_proj.weight"])
    dtq, _, dtv = mha_bwd(w, g, req, "self_attn.", dctx1, sv1, nhead, B, need=(True, True, need_tgt),
                          drop_site=_site(drop, 0))
    d_qpos = K.add_bf16(dcq, dtq)
    d_tgt = K.add_bf16(ds1, dtq, dtv) if need_tgt else None
    return d_tgt, d_qpos, dmem_pos, dmem


# ------------------------------------------------------------------------------------------------ RoBERTa layer
def roberta_layer_fwd(w: W, x, key_mask, nhead: int, B: int, eps: float, drop: Optional[Drop] = None):
    """transformers RobertaLayer (post-LN BERT block, erf GELU).  x [L*B, E] bf16, rows l*B + b.  Dropout sites:
    0 attention_probs_dropout, 1 attention.output.dropout, 2 output.dropout (hidden_dropout_prob)."""
    M, E = x.shape
    L = M // B
    Wqkv = w["attention.self.qkv"]
    qkv = torch.empty((M, 3 * E), dtype=BF, device=x.device)
    for i, nm in enumerate(("query", "key", "value")):
        K.linear_fwd(x, Wqkv[i * E:(i + 1) * E], w[f"attention.self.{nm}.bias"], out=qkv[:, i * E:(i + 1) * E])
    q3, k3, v3 = (qkv[:, i * E:(i + 1) * E].view(L, B, E) for i in range(3))
    ctx, probs = K.attention_fwd(q3, k3, v3, key_mask, nhead, drop=_site(drop, 0))
    ctx = ctx.view(M, E)
    x1, _, s1, m1, r1 = branch_ln_fwd(w, ctx, "attention.output.dense.weight", "attention.output.dense.bias", x,
                                      _site(drop, 1), "attention.output.LayerNorm.", eps)
    pre = torch.empty((M, w["intermediate.dense.weight"].shape[0]), dtype=BF, device=x.device)
    h = K.linear_fwd(x1, w["intermediate.dense.weight"], w["intermediate.dense.bias"], act=ACT_GELU, aux=pre)
    x2, _, s2, m2, r2 = branch_ln_fwd(w, h, "output.dense.weight", "output.dense.bias", x1, _site(drop, 2),
                                      "output.LayerNorm.", eps)
    return x2, (x, qkv, probs, ctx, s1, m1, r1, x1, pre, h, s2, m2, r2)


def roberta_layer_bwd(w: W, g: G, req: Set[str], dy, saved, nhead: int, B: int, need_dx: bool = True,
                      drop: Optional[Drop] = None):
    x, qkv, probs, ctx, s1, m1, r1, x1, pre, h, s2, m2, r2 = saved
    M, E = x.shape
    L = M // B
    ds2, dy2 = branch_ln_bwd(w, g, req, "output.LayerNorm.", dy, s2, m2, r2, _site(drop, 2))
    lin_param_grads(g, req, "output.dense.weight", "output.dense.bias", dy2, h, w["output.dense.weight"].shape)
    dh = K.linear_dgrad(dy2, w["output.dense.weight"])
    dpre = K.gelu_bwd(dh, pre)
    lin_param_grads(g, req, "intermediate.dense.weight", "intermediate.dense.bias", dpre, x1,
                    w["intermediate.dense.weight"].shape)
    dx1 = K.linear_dgrad(dpre, w["intermediate.dense.weight"], res=ds2)
    ds1, dy1 = branch_ln_bwd(w, g, req, "attention.output.LayerNorm.", dx1, s1, m1, r1, _site(drop, 1))
    lin_param_grads(g, req, "attention.output.dense.weight", "attention.output.dense.bias", dy1, ctx,
                    w["attention.output.dense.weight"].shape)
    dctx = K.linear_dgrad(dy1, w["attention.output.dense.weight"])
    dqkv = torch.empty((M, 3 * E), dtype=BF, device=x.device)
    q3, k3, v3 = (qkv[:, i * E:(i + 1) * E].view(L, B, E) for i in range(3))
    d3 = [dqkv[:, i * E:(i + 1) * E].view(L, B, E) for i in range(3)]
    K.attention_bwd(dctx.view(L, B, E), q3, k3, v3, probs, nhead, d3[0], d3[1], d3[2], drop=_site(drop, 0))
    for i, nm in enumerate(("query", "key", "value")):
        lin_param_grads(g, req, f"attention.self.{nm}.weight", f"attention.self.{nm}.bias",
                        dqkv[:, i * E:(i + 1) * E], x, (E, E))
    if not need_dx:
        return None
    return K.linear_dgrad(dqkv, w["attention.self.qkv"], res=ds1)


# ------------------------------------------------------------------------------------------------ bottleneck
def bottleneck_fwd(w: W, x, stride: int, has_ds: bool, into=None):
    """torchvision Bottleneck (v1.5: stride on the 3x3) with FrozenBatchNorm folded: conv weights carry the BN scale,
    the epilogue adds the BN shift (models/backbone.py:48-58).  x NHWC bf16.  `into` = (y1, y2, out) buffers to write
    (batch slices of full-batch tensors when the trunk runs as two half-batch chains)."""
    b1, b2, bo = into if into is not None else (None, None, None)
    y1 = K.conv_fwd(x, w["conv1.weight"], w["bn1.shift"], act=ACT_RELU, out=b1)
    y2 = K.conv_fwd(y1, w["conv2.weight"], w["bn2.shift"], stride=stride, pad=1, act=ACT_RELU, out=b2)
    idt = K.conv_fwd(x, w["downsample.0.weight"], w["downsample.1.shift"], stride=stride) if has_ds else x
    out = K.conv_fwd(y2, w["conv3.weight"], w["bn3.shift"], res=idt, act=ACT_RELU, out=bo)
    return out, (x, y1, y2)


def _conv_wgrad_param(g: G, name: str, dy, x, w_shadow, scale, stride: int, pad: int) -> None:
    cout, kh, kw, cin = w_shadow.shape
    with K.wgrad_lane(dy, x):
        if kh * kw > 1:
            # the GEMM produces [Cout, taps, Cin]; the parameter's layout is [Cout, Cin, taps].  The GEMM accumulates
            # (split-K) into a zeroed scratch OUTSIDE the gradient arena and the permute writes the final layout INTO
            # the arena: every gradient of the stage then travels in the stage's one flat all-reduce (a gradient
            # outside the arena costs a collective of its own: ~30 trailing all-reduces per step before this)
            tmp = torch.zeros((cout, kh, kw, cin), dtype=torch.float32, device=dy.device)
            K.conv_wgrad(dy, x, tmp, stride=stride, pad=pad, row_scale=scale)
            dw = _zeros((cout, cin, kh, kw), dy.device)
            K.permute_021(tmp.view(cout, kh * kw, cin), out=dw.view(cout, cin, kh * kw))
        else:
            dw = _zeros((cout, kh, kw, cin), dy.device)
            K.conv_wgrad(dy, x, dw, stride=stride, pad=pad, row_scale=scale)
    g[name] = dw.view(cout, cin, kh, kw)


def bottleneck_bwd_chains(w: W, g: G, req: Set[str], gz, saved, stride: int, has_ds: bool, need_dx: bool, streams):
    """`bottleneck_bwd` with the data-gradient chain split into len(streams) independent batch slices, one stream each
    (runtime.backbone_bwd); the weight gradients stay full-batch launches on the weight-gradient lane, which waits for
    every chain before each of them (kernels._WgradLane.also_wait).  Same arithmetic per image."""
    x, y1, y2 = saved
    n = x.shape[0]
    k = len(streams)
    per = n // k
    sls = [slice(i * per, (i + 1) * per) for i in range(k)]
    g2 = torch.empty_like(y2)
    g1 = torch.empty_like(y1)
    dx = torch.empty_like(x) if need_dx else None
    K.keep_alive(gz, g2, g1, dx)  # made on the main stream, read by the other chain's stream until the stage's join
    if "conv3.weight" in req:
        _conv_wgrad_param(g, "conv3.weight", gz, y2, w["conv3.weight"], w["bn3.scale"], 1, 0)
    with K.gemm_chains(k):
        for s_, sl in zip(streams, sls):
            with torch.cuda.stream(s_):
                K.conv_dgrad(gz[sl], w["conv3.weight"], y2.shape[1:3], mask=y2[sl], out=g2[sl])
    if "conv2.weight" in req:
        _conv_wgrad_param(g, "conv2.weight", g2, y1, w["conv2.weight"], w["bn2.scale"], stride, 1)
    with K.gemm_chains(k):
        for s_, sl in zip(streams, sls):
            with torch.cuda.stream(s_):
                K.conv_dgrad(g2[sl], w["conv2.weight"], y1.shape[1:3], stride=stride, pad=1, mask=y1[sl], out=g1[sl])
    if "conv1.weight" in req:
        _conv_wgrad_param(g, "conv1.weight", g1, x, w["conv1.weight"], w["bn1.scale"], 1, 0)
    if has_ds and "downsample.0.weight" in req:
        _conv_wgrad_param(g, "downsample.0.weight", gz, x, w["downsample.0.weight"], w["downsample.1.scale"], stride, 0)
    if not need_dx:
        return None
    with K.gemm_chains(k):
        for s_, sl in zip(streams, sls):
            with torch.cuda.stream(s_):
                gi = K.conv_dgrad(gz[sl], w["downsample.0.weight"], x.shape[1:3], stride=stride) if has_ds else gz[sl]
                K.conv_dgrad(g1[sl], w["conv1.weight"], x.shape[1:3], res=gi, mask=x[sl], out=dx[sl])
    return dx


def bottleneck_bwd(w: W, g: G, req: Set[str], gz, saved, stride: int, has_ds: bool, need_dx: bool):
    """gz: gradient w.r.t. the pre-ReLU sum (already multiplied by out > 0).  Returns the gradient w.r.t. the block
    input *already masked by (x > 0)*, i.e. the `gz` of the preceding block."""
    x, y1, y2 = saved
    if "conv3.weight" in req:
        _conv_wgrad_param(g, "conv3.weight", gz, y2, w["conv3.weight"], w["bn3.scale"], 1, 0)
    g2 = K.conv_dgrad(gz, w["conv3.weight"], y2.shape[1:3], mask=y2)
    if "conv2.weight" in req:
        _conv_wgrad_param(g, "conv2.weight", g2, y1, w["conv2.weight"], w["bn2.scale"], stride, 1)
    g1 = K.conv_dgrad(g2, w["conv2.weight"], y1.shape[1:3], stride=stride, pad=1, mask=y1)
    if "conv1.weight" in req:
        _conv_wgrad_param(g, "conv1.weight", g1, x, w["conv1.weight"], w["bn1.scale"], 1, 0)
    if has_ds and "downsample.0.weight" in req:
        _conv_wgrad_param(g, "downsample.0.weight", gz, x, w["downsample.0.weight"], w["downsample.1.scale"], stride, 0)
    if not need_dx:
        return None
    gi = K.conv_dgrad(gz, w["downsample.0.weight"], x.shape[1:3], stride=stride) if has_ds else gz
    return K.conv_dgrad(g1, w["conv1.weight"], x.shape[1:3], res=gi, mask=x)


# ------------------------------------------------------------------------------------------------ strided GEMM helpers
def seq_from_nhwc_fwd(x, wt, bias, out, B: int):
    """input_proj (models/mdetr.py:351,383) fused with flatten(2).permute(2,0,1) (transformer.py:101):
    x NHWC [B,H,W,Cin] bf16 -> out rows (y*W + x)*B + b, [H*W*B, Cout] (a row slice of the encoder source)."""
    n, h, wd, cin = x.shape
    cout = wt.shape[0]
    K.gemm(GEMM_FWD, K._nhwc_t4(x), K.t4(wt, (cin, cout, 1, 1), (1, cin, 0, 0)), out, ext=(wd, h, n),
           tile=K.pick_tile(wd, h, n), n_cols=cout, out_strides=(B * cout, wd * B * cout, cout), k_per_tap=cin,
           col_shift=bias)
    return out


def _seq_as_nhwc_t4(d, h: int, wd: int, B: int):
    """[H*W*B, C] sequence rows viewed as the (c, x, y, n) pixel space of an NHWC tensor."""
    c = d.shape[1]
    ld = d.stride(0)
    return K.t4(d, (c, wd, h, B), (1, B * ld, wd * B * ld, ld))


def seq_from_nhwc_bwd(g: G, req: Set[str], wname: str, bname: str, dseq, x, wt, need_dx: bool):
    n, h, wd, cin = x.shape
    cout = wt.shape[0]
    a = _seq_as_nhwc_t4(dseq, h, wd, n)
    if wname in req:
        dw = _zeros((cout, cin), x.device)
        tile = K.pick_tile(wd, h, n, 64, 64)
        ptiles = -(-wd // tile[0]) * -(-h // tile[1]) * -(-n // tile[2])
        K.gemm(GEMM_WGRAD, a, K._nhwc_t4(x), dw, ext=(wd, h, n), tile=tile, n_cols=cin, m_rows=cout,
               out_strides=(cin, 0, 0), splits=K._wgrad_splits(cout, cin, 1, ptiles), accumulate=True)
        g[wname] = dw.view(cout, cin, 1, 1)
    if bname in req:
        db = _zeros((cout,), x.device)
        K.colsum(dseq, db)
        g[bname] = db
    if not need_dx:
        return None
    dx = torch.empty_like(x)
    K.gemm(GEMM_DGRAD, a, K.t4(wt, (cin, cout, 1, 1), (1, cin, 0, 0)), dx, ext=(wd, h, n), tile=K.pick_tile(wd, h, n),
           n_cols=cin, out_strides=(cin, wd * cin, h * wd * cin), k_per_tap=cout)
    return dx


def heads_linear_fwd(hs, wt, bias, L: int, Q: int, B: int, act: int = ACT_NONE):
    """Linear over decoder states hs [L, Q*B, E] (rows q*B + b) writing fp32 [L, B, Q, N]: the hs.transpose(1, 2) of
    models/transformer.py:188 is folded into the output addressing."""
    E = hs.shape[-1]
    N = wt.shape[0]
    out = torch.empty((L, B, Q, N), dtype=torch.float32, device=hs.device)
    K.gemm(GEMM_FWD, K.t4(hs, (E, B, Q, L), (1, E, B * E, Q * B * E)), K.t4(wt, (E, N, 1, 1), (1, E, 0, 0)), out,
           ext=(B, Q, L), tile=K.pick_tile(B, Q, L), n_cols=N, out_strides=(Q * N, N, B * Q * N), k_per_tap=E,
           col_shift=bias, act=act)
    return out


def heads_linear_bwd(g: G, req: Set[str], wname: str, bname: str, dout16, hs, wt, L: int, Q: int, B: int,
                     res=None, mask=None):
    """dout16 bf16 [L, B, Q, ldn] (ldn >= N, zero padded to a multiple of 8) -> d_hs bf16 [L, Q*B, E]
    ((+ res) * (mask > 0)), parameter grads into g."""
    E = hs.shape[-1]
    N = wt.shape[0]
    ldn = dout16.shape[-1]
    a = K.t4(dout16, (ldn, B, Q, L), (1, Q * ldn, ldn, B * Q * ldn))
    if wname in req:
        dw = _zeros((N, E), hs.device)
        tile = K.pick_tile(B, Q, L, 64, 64)
        ptiles = -(-B // tile[0]) * -(-Q // tile[1]) * -(-L // tile[2])
        K.gemm(GEMM_WGRAD, a, K.t4(hs, (E, B, Q, L), (1, E, B * E, Q * B * E)), dw, ext=(B, Q, L), tile=tile,
               n_cols=E, m_rows=N, out_strides=(E, 0, 0), splits=K._wgrad_splits(N, E, 1, ptiles), accumulate=True)
        g[wname] = dw
    if bname in req:
        db = _zeros((N,), hs.device)
        K.colsum(dout16.view(-1, ldn)[:, :N], db)
        g[bname] = db
    dhs = torch.empty((L, Q * B, E), dtype=BF, device=hs.device)
    K.gemm(GEMM_DGRAD, a, K.t4(wt, (E, N, 1, 1), (1, E, 0, 0)), dhs, ext=(B, Q, L), tile=K.pick_tile(B, Q, L),
           n_cols=E, out_strides=(E, B * E, Q * B * E), k_per_tap=ldn, res=res, mask=mask)
    return dhs
