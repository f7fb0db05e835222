"""Mask branch of DETRsegm as one autograd stage (reference models/segmentation.py:157-167, 170-273).

forward : hs[-1] --q_linear--> q, encoder image memory --k_linear--> k, per-head softmax over pixels (MHAttentionMap),
          concat with the projected image features, 5 x (3x3 conv + GroupNorm(8) + ReLU) with nearest upsampling and
          FPN adapters in between, 3x3 conv to one channel  ->  pred_masks [B, Q, H/4, W/4] fp32.
backward: hand written; gradients for the mask-branch parameters and, when the detector is not frozen, for hs,
          the encoder memory, src_proj and the three backbone feature maps.

All maps are NHWC bf16 with B*Q maps per batch; the convolutions run on the implicit-GEMM engine, GroupNorm / upsample /
input assembly on the kernels in csrc/maskhead.cu.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import blocks as Bk
from . import kernels as K
from .runtime import Call, GView, RView, Spec, WView

BF = torch.bfloat16
_LAYERS = (("lay1", "gn1"), ("lay2", "gn2"), ("lay3", "gn3"), ("lay4", "gn4"), ("lay5", "gn5"))


def _w11(wt):
    """[Cout, Cin] shadow of a 1x1 convolution as the [Cout, 1, 1, Cin] filter the conv wrappers expect."""
    return wt.view(wt.shape[0], 1, 1, wt.shape[1])


def _conv_gn(w: WView, lay: str, gn: str, x):
    z = K.conv_fwd(x, w[f"mask_head.{lay}.weight"], w[f"mask_head.{lay}.bias"], pad=1)
    a, m, r = K.groupnorm_relu_fwd(z, w[f"mask_head.{gn}.weight"], w[f"mask_head.{gn}.bias"])
    return a, (x, z, m, r)


def _conv_param_grads(g, rq, name: str, dz, x, w_shadow, pad: int) -> None:
    """Weight (OIHW, fp32) and bias gradients of a biased convolution; dz may have zero-padded channels."""
    cout, kh, kw, cin = w_shadow.shape
    if (name + ".weight") in rq:
        if kh * kw > 1:  # scratch outside the arena, final layout inside it (see blocks._conv_wgrad_param)
            tmp = torch.zeros((cout, kh, kw, cin), dtype=torch.float32, device=dz.device)
            K.conv_wgrad(dz, x, tmp, pad=pad)
            dw = Bk._zeros((cout, cin, kh, kw), dz.device)
            K.permute_021(tmp.view(cout, kh * kw, cin), out=dw.view(cout, cin, kh * kw))
        else:
            dw = Bk._zeros((cout, kh, kw, cin), dz.device)
            K.conv_wgrad(dz, x, dw, pad=pad)
        g[name + ".weight"] = dw.view(cout, cin, kh, kw)
    if (name + ".bias") in rq:
        c = dz.shape[-1]
        db = Bk._zeros((c,), dz.device)
        K.colsum(dz.view(-1, c), db)
        g[name + ".bias"] = db[:cout]


def _conv_gn_bwd(w: WView, g, rq, lay: str, gn: str, da, saved, need_dx: bool = True):
    x, z, m, r = saved
    gw, gb = f"mask_head.{gn}.weight", f"mask_head.{gn}.bias"
    dg = db = None
    if gw in rq:
        dg = Bk._zeros((z.shape[-1],), z.device)
        db = Bk._zeros((z.shape[-1],), z.device)
        g[gw], g[gb] = dg, db
    dz = K.groupnorm_relu_bwd(da, z, m, r, w[gw], w[gb], dg, db)
    ws = w[f"mask_head.{lay}.weight"]
    _conv_param_grads(g, rq, f"mask_head.{lay}", dz, x, ws, 1)
    return K.conv_dgrad(dz, ws, x.shape[1:3], pad=1) if need_dx else None


def mask_fwd(c: Call, hs, mem32, src_proj, c4, c3, c2, small_mask):
    """hs bf16 [L, Q*B, E]; mem32 fp32 [S, B, E]; src_proj bf16 [hw, B, E]; c4/c3/c2 NHWC bf16 backbone maps
    (strides 16 / 8 / 4); small_mask uint8 [B, h, w].  Returns pred_masks fp32 [B, Q, H1, W1]."""
    st = c.stage
    w = WView(c.w, "")
    B, h, wd = small_mask.shape
    hw = h * wd
    E = hs.shape[-1]
    Q = hs.shape[1] // B
    NH = st.nheads
    # ---- MHAttentionMap (segmentation.py:262-273)
    q_in = hs[-1]                                             # [Q*B, E] rows q*B + b
    k_in = K.cast_bf16(mem32[:hw].contiguous().view(hw * B, E))  # image rows of the encoder output
    qp = K.linear_fwd(q_in, w["bbox_attention.q_linear.weight"], w["bbox_attention.q_linear.bias"])
    kp = K.linear_fwd(k_in, w["bbox_attention.k_linear.weight"], w["bbox_attention.k_linear.bias"])
    probs = K.attn_map_fwd(qp.view(Q, B, E), kp.view(hw, B, E), small_mask.view(B, hw), NH)
    # ---- MaskHeadSmallConv (segmentation.py:203-241)
    x0 = K.mask_input(src_proj, probs, B, Q, h, wd)
    a1, s1 = _conv_gn(w, "lay1", "gn1", x0)
    a2, s2 = _conv_gn(w, "lay2", "gn2", a1)
    f1 = K.conv_fwd(c4, _w11(w["mask_head.adapter1.weight"]), w["mask_head.adapter1.bias"])
    u3 = K.upsample_add(a2, f1, Q)
    a3, s3 = _conv_gn(w, "lay3", "gn3", u3)
    f2 = K.conv_fwd(c3, _w11(w["mask_head.adapter2.weight"]), w["mask_head.adapter2.bias"])
    u4 = K.upsample_add(a3, f2, Q)
    a4, s4 = _conv_gn(w, "lay4", "gn4", u4)
    f3 = K.conv_fwd(c2, _w11(w["mask_head.adapter3.weight"]), w["mask_head.adapter3.bias"])
    u5 = K.upsample_add(a4, f3, Q)
    a5, s5 = _conv_gn(w, "lay5", "gn5", u5)
    out = K.conv_fwd(a5, w["mask_head.out_lay.weight"], w["mask_head.out_lay.bias"], pad=1, out_dtype=torch.float32)
    H1, W1 = out.shape[1:3]
    saved = None
    if c.save:
        saved = (q_in, k_in, qp, kp, probs, (B, Q, h, wd, E, NH), s1, s2, s3, s4, s5, a5, c4, c3, c2,
                 (a2.shape[1:3], a3.shape[1:3], a4.shape[1:3]))
    return (out.view(B, Q, H1, W1),), saved


def mask_bwd(c: Call, saved, needs, dpred):
    q_in, k_in, qp, kp, probs, (B, Q, h, wd, E, NH), s1, s2, s3, s4, s5, a5, c4, c3, c2, small_hw = saved
    w = WView(c.w, "")
    grads: Dict[str, torch.Tensor] = {}
    g, rq = GView(grads, ""), RView(c.req, "")
    need_hs, need_mem, need_src, need_c4, need_c3, need_c2 = needs[0], needs[1], needs[2], needs[3], needs[4], needs[5]
    N = B * Q
    H1, W1 = dpred.shape[-2:]
    d16 = K.cast_pad_bf16(dpred.contiguous().view(N, H1, W1, 1), 8)
    ws = w["mask_head.out_lay.weight"]
    _conv_param_grads(g, rq, "mask_head.out_lay", d16, a5, ws, 1)
    da = K.conv_dgrad(d16, ws, (H1, W1), pad=1)
    fpn_in = {"adapter3": c2, "adapter2": c3, "adapter1": c4}
    need_fpn = {"adapter3": need_c2, "adapter2": need_c3, "adapter1": need_c4}
    dfeat = {}
    for (lay, gn), sv, ad, shw in ((("lay5", "gn5"), s5, "adapter3", small_hw[2]), (("lay4", "gn4"), s4, "adapter2", small_hw[1]),
                                   (("lay3", "gn3"), s3, "adapter1", small_hw[0])):
        du = _conv_gn_bwd(w, g, rq, lay, gn, da, sv)
        want_f = (f"mask_head.{ad}.weight" in rq) or need_fpn[ad]
        da, df = K.upsample_add_bwd(du, tuple(shw), Q, want_f)
        if want_f:
            wa = w[f"mask_head.{ad}.weight"]  # [Cout, Cin] shadow of the 1x1 adapter
            cout, cin = wa.shape
            wa4 = wa.view(cout, 1, 1, cin)
            x = fpn_in[ad]
            if (f"mask_head.{ad}.weight") in rq:
                dw = Bk._zeros((cout, 1, 1, cin), x.device)
                K.conv_wgrad(df, x, dw)
                g[f"mask_head.{ad}.weight"] = dw.view(cout, cin, 1, 1)
            if (f"mask_head.{ad}.bias") in rq:
                db = Bk._zeros((cout,), x.device)
                K.colsum(df.view(-1, cout), db)
                g[f"mask_head.{ad}.bias"] = db
            if need_fpn[ad]:
                dfeat[ad] = K.conv_dgrad(df, wa4, x.shape[1:3])
    da = _conv_gn_bwd(w, g, rq, "lay2", "gn2", da, s2)
    attn_req = any(n.startswith("bbox_attention.") for n in c.req)
    need_x0 = need_src or need_hs or need_mem or attn_req
    dx0 = _conv_gn_bwd(w, g, rq, "lay1", "gn1", da, s1, need_dx=need_x0)
    d_hs = d_mem = d_src = None
    if need_x0:
        d_src, dprobs = K.mask_input_bwd(dx0, B, Q, E, NH, probs.shape[-1], need_src)
        hw = h * wd
        dqp = torch.empty_like(qp)
        dkp = torch.empty_like(kp)
        K.attn_map_bwd(dprobs, qp.view(Q, B, E), kp.view(hw, B, E), probs, NH, dqp.view(Q, B, E), dkp.view(hw, B, E))
        Bk.lin_param_grads(g, rq, "bbox_attention.q_linear.weight", "bbox_attention.q_linear.bias", dqp, q_in, (E, E))
        Bk.lin_param_grads(g, rq, "bbox_attention.k_linear.weight", "bbox_attention.k_linear.bias", dkp, k_in, (E, E))
        if need_hs:
            L = c.n_dec_layers
            d_hs = torch.zeros((L, Q * B, E), dtype=BF, device=dx0.device)
            K.linear_dgrad(dqp, w["bbox_attention.q_linear.weight"], out=d_hs[-1])
        if need_mem:
            S = c.seq_len
            d_mem = torch.zeros((S, B, E), dtype=torch.float32, device=dx0.device)
            dk_in = K.linear_dgrad(dkp, w["bbox_attention.k_linear.weight"])
            K.cast_f32(dk_in, out=d_mem[:hw].view(hw * B, E))
    return (d_hs, d_mem, d_src, dfeat.get("adapter1"), dfeat.get("adapter2"), dfeat.get("adapter3"), None), grads


MASKHEAD = Spec("maskhead", 7, mask_fwd, mask_bwd)
