"""Block-level forward AND backward parity on the GPU: each hand-written block (blocks.py / runtime.py) against the
oracle's fp32 torch restatement of the same reference code, on identical bf16-representable inputs and weights.

Tolerance: the blocks keep their intermediate activations and gradients in bf16 (unit round-off 2**-8 = 3.9e-3).
Forward outputs land within a few 1e-3 norm-wise (assert 1e-2).  Backward through a ReLU is noisier: the stored
activation differs from the fp32 one by ~1e-3 relative, which flips the ReLU mask on a fraction p ~ 1e-3 of the
elements, an O(1) error on those elements, i.e. ~sqrt(p) ~ 3e-2 norm-wise (measured 2.5e-2 .. 3.4e-2 on B200); blocks
containing a ReLU therefore assert 8e-2 and the bottleneck is additionally checked at 1e-2 against a reference that
rounds its stored activations like we do, where no mask can flip.  A wrong formula, a transposed operand or a missing
term shows up as an error of order 1.
"""
from __future__ import annotations

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
DEV = "cuda"
FWD_TOL, BWD_TOL, TIGHT = 1e-2, 8e-2, 1e-2


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference_math():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def q(t):
    """Round to bf16-representable fp32 so the bf16 shadow equals the master weight exactly."""
    return t.to(BF).float()


def rnd(*shape, scale=1.0):
    return q(torch.randn(*shape, device=DEV) * scale)


def split_w(sd):
    """name -> tensor the kernels read: bf16 for matrices, fp32 for vectors (what ShadowBank provides)."""
    return {k: (v.to(BF).contiguous() if v.dim() >= 2 else v.contiguous()) for k, v in sd.items()}


def leaf(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


def check_grads(g, ref_sd, skip=(), tol=BWD_TOL):
    bad = []
    for k, p in ref_sd.items():
        if k in skip or p.grad is None:
            continue
        assert k in g, f"missing gradient {k}"
        e = rel_err(g[k].reshape(p.shape), p.grad)
        if not e <= tol:
            bad.append((k, e))
    assert not bad, bad


def mha_sd(prefix, E):
    return {prefix + "in_proj_weight": rnd(3 * E, E, scale=E ** -0.5), prefix + "in_proj_bias": rnd(3 * E, scale=0.1),
            prefix + "out_proj.weight": rnd(E, E, scale=E ** -0.5), prefix + "out_proj.bias": rnd(E, scale=0.1)}


def ln_sd(prefix, E):
    return {prefix + "weight": rnd(E, scale=0.2) + 1, prefix + "bias": rnd(E, scale=0.1)}


def ffn_sd(E, F):
    return {"linear1.weight": rnd(F, E, scale=E ** -0.5), "linear1.bias": rnd(F, scale=0.1),
            "linear2.weight": rnd(E, F, scale=F ** -0.5), "linear2.bias": rnd(E, scale=0.1)}


def test_encoder_layer():
    from oracle import model as O
    from toist_b200 import blocks as Bk
    from toist_b200 import kernels as K

    torch.manual_seed(0)
    E, S, B, H = 256, 61, 2, 8
    sd = {**mha_sd("self_attn.", E), **ffn_sd(E, 2048), **ln_sd("norm1.", E), **ln_sd("norm2.", E)}
    x, pos = rnd(S * B, E), rnd(S * B, E)
    km = torch.zeros(B, S, dtype=torch.uint8, device=DEV)
    km[1, S - 5:] = 1
    g = {}
    y, saved, _ = Bk.encoder_layer_fwd(split_w(sd), x.to(BF), pos.to(BF), km, H, B)
    dy = rnd(S * B, E)
    dx = K.add_bf16(*Bk.encoder_layer_bwd(split_w(sd), g, set(sd), dy.to(BF), saved, H, B))
    ref = leaf({"layers.0." + k: v for k, v in sd.items()})
    xr = x.view(S, B, E).clone().requires_grad_(True)
    yr = O.encoder(xr, km.bool(), pos.view(S, B, E), ref, "", 1, H)
    yr.backward(dy.view(S, B, E))
    assert rel_err(y.float().view(S, B, E), yr) < FWD_TOL
    assert rel_err(dx.float().view(S, B, E), xr.grad) < BWD_TOL
    check_grads({"layers.0." + k: v for k, v in g.items()}, ref)


def test_decoder_layer():
    from oracle import model as O
    from toist_b200 import blocks as Bk
    from toist_b200 import kernels as K

    torch.manual_seed(1)
    E, S, Q, B, H = 256, 77, 20, 2, 8
    sd = {**mha_sd("self_attn.", E), **mha_sd("cross_attn_image.", E), **ffn_sd(E, 2048), **ln_sd("norm1.", E),
          **ln_sd("norm3.", E), **ln_sd("norm4.", E)}
    tgt, qpos, mem, pos = rnd(Q * B, E), rnd(Q * B, E), rnd(S * B, E), rnd(S * B, E)
    km = torch.zeros(B, S, dtype=torch.uint8, device=DEV)
    km[0, S - 9:] = 1
    w = split_w(sd)
    g = {}
    mem_pos = K.add_bf16(mem.to(BF), pos.to(BF))
    y, saved, _ = Bk.decoder_layer_fwd(w, tgt.to(BF), qpos.to(BF), mem.to(BF), mem_pos, km, H, B)
    dy = rnd(Q * B, E)
    d_tgt, d_qpos, d_mp, d_mem = Bk.decoder_layer_bwd(w, g, set(sd), dy.to(BF), None, saved, H, B)
    ref = leaf({"layers.0." + k: v for k, v in sd.items()})
    tr = tgt.view(Q, B, E).clone().requires_grad_(True)
    qr = qpos.view(Q, B, E).clone().requires_grad_(True)
    mr = mem.view(S, B, E).clone().requires_grad_(True)
    pr = pos.view(S, B, E).clone().requires_grad_(True)
    # one layer of oracle.decoder without the shared final norm (that norm lives in runtime.decoder_fwd)
    x = tr
    p = "layers.0."
    qk = x + qr
    x = O._ln(x + O.mha(qk, qk, x, ref, p + "self_attn.", H, None), ref, p + "norm1.")
    x = O._ln(x + O.mha(x + qr, mr + pr, mr, ref, p + "cross_attn_image.", H, km.bool()), ref, p + "norm3.")
    yr = O._ln(x + O._ffn(x, ref, p), ref, p + "norm4.")
    yr.backward(dy.view(Q, B, E))
    assert rel_err(y.float().view(Q, B, E), yr) < FWD_TOL
    assert rel_err(d_tgt.float().view(Q, B, E), tr.grad) < BWD_TOL
    assert rel_err(d_qpos.float().view(Q, B, E), qr.grad) < BWD_TOL
    # memory feeds the keys through (memory + pos) and the values directly
    assert rel_err((d_mem.float() + d_mp.float()).view(S, B, E), mr.grad) < BWD_TOL
    assert rel_err(d_mp.float().view(S, B, E), pr.grad) < BWD_TOL
    check_grads({"layers.0." + k: v for k, v in g.items()}, ref)


def test_roberta_layer():
    from oracle import model as O
    from toist_b200 import blocks as Bk

    torch.manual_seed(2)
    E, L, B, H, F = 768, 12, 3, 12, 3072
    p = "encoder.layer.0."
    sd = {}
    for nm in ("query", "key", "value"):
        sd[f"attention.self.{nm}.weight"] = rnd(E, E, scale=E ** -0.5)
        sd[f"attention.self.{nm}.bias"] = rnd(E, scale=0.1)
    sd.update({"attention.output.dense.weight": rnd(E, E, scale=E ** -0.5), "attention.output.dense.bias": rnd(E, scale=0.1),
               "intermediate.dense.weight": rnd(F, E, scale=E ** -0.5), "intermediate.dense.bias": rnd(F, scale=0.1),
               "output.dense.weight": rnd(E, F, scale=F ** -0.5), "output.dense.bias": rnd(E, scale=0.1),
               **ln_sd("attention.output.LayerNorm.", E), **ln_sd("output.LayerNorm.", E)})
    w = split_w(sd)
    w["attention.self.qkv"] = torch.cat([w[f"attention.self.{nm}.weight"] for nm in ("query", "key", "value")], 0)
    x = rnd(L * B, E)  # rows l*B + b
    attn = torch.ones(B, L, dtype=torch.int64, device=DEV)
    attn[1, 8:] = 0
    km = attn.ne(1).to(torch.uint8)
    g = {}
    y, saved = Bk.roberta_layer_fwd(w, x.to(BF), km, H, B, 1e-5)
    dy = rnd(L * B, E)
    dy3 = dy.view(L, B, E).clone()
    dx = Bk.roberta_layer_bwd(w, g, set(sd), dy3.view(L * B, E).to(BF), saved, H, B)
    # reference: one RoBERTa layer taken out of oracle.roberta_encode (batch-first there)
    ref = leaf({p + k: v for k, v in sd.items()})
    xr = x.view(L, B, E).transpose(0, 1).contiguous().requires_grad_(True)  # [B, L, E]
    dh = E // H
    bias = torch.zeros(B, 1, 1, L, device=DEV).masked_fill(attn[:, None, None, :] == 0, torch.finfo(torch.float32).min)
    import math
    import torch.nn.functional as F_

    qh = F_.linear(xr, ref[p + "attention.self.query.weight"], ref[p + "attention.self.query.bias"])
    kh = F_.linear(xr, ref[p + "attention.self.key.weight"], ref[p + "attention.self.key.bias"])
    vh = F_.linear(xr, ref[p + "attention.self.value.weight"], ref[p + "attention.self.value.bias"])
    qh, kh, vh = (t.view(B, L, H, dh).transpose(1, 2) for t in (qh, kh, vh))
    att = (qh @ kh.transpose(-1, -2)) / math.sqrt(dh) + bias
    ctx = (att.softmax(-1) @ vh).transpose(1, 2).reshape(B, L, E)
    yy = F_.linear(ctx, ref[p + "attention.output.dense.weight"], ref[p + "attention.output.dense.bias"])
    x1 = F_.layer_norm(yy + xr, (E,), ref[p + "attention.output.LayerNorm.weight"], ref[p + "attention.output.LayerNorm.bias"], 1e-5)
    hh = F_.gelu(F_.linear(x1, ref[p + "intermediate.dense.weight"], ref[p + "intermediate.dense.bias"]))
    yy = F_.linear(hh, ref[p + "output.dense.weight"], ref[p + "output.dense.bias"])
    yr = F_.layer_norm(yy + x1, (E,), ref[p + "output.LayerNorm.weight"], ref[p + "output.LayerNorm.bias"], 1e-5)
    yr.backward(dy3.transpose(0, 1))
    assert rel_err(y.float().view(L, B, E).transpose(0, 1), yr) < FWD_TOL
    assert rel_err(dx.float().view(L, B, E).transpose(0, 1), xr.grad) < BWD_TOL
    # d(key bias) is analytically zero (softmax is invariant to a per-query shift): compare absolutely
    check_grads({p + k: v for k, v in g.items()}, ref, skip=(p + "attention.self.key.bias",))
    kb = g["attention.self.key.bias"]
    assert float(kb.abs().max()) < 2e-2 * float(g["attention.self.query.bias"].abs().max() + 1e-6)


@pytest.mark.parametrize("cin,planes,stride,ds,hw", [(256, 128, 2, True, 24), (512, 128, 1, False, 12),
                                                     (1024, 512, 2, True, 10), (64, 64, 1, True, 20)])
def test_bottleneck(cin, planes, stride, ds, hw):
    import torch.nn.functional as F_

    from toist_b200 import blocks as Bk

    torch.manual_seed(3)
    N = 2
    cout = planes * 4
    convs = {"conv1": rnd(planes, cin, 1, 1, scale=(cin) ** -0.5), "conv2": rnd(planes, planes, 3, 3, scale=(9 * planes) ** -0.5),
             "conv3": rnd(cout, planes, 1, 1, scale=planes ** -0.5)}
    bns = {"bn1": planes, "bn2": planes, "bn3": cout}
    if ds:
        convs["downsample.0"] = rnd(cout, cin, 1, 1, scale=cin ** -0.5)
        bns["downsample.1"] = cout
    pair = {"conv1": "bn1", "conv2": "bn2", "conv3": "bn3", "downsample.0": "downsample.1"}
    scale = {k: q(torch.rand(c, device=DEV) + 0.5) for k, c in bns.items()}
    shift = {k: rnd(c, scale=0.1) for k, c in bns.items()}
    w = {}
    for cn, cw in convs.items():
        bn = pair[cn]
        w[cn + ".weight"] = (cw * scale[bn][:, None, None, None]).permute(0, 2, 3, 1).contiguous().to(BF)
        w[bn + ".scale"] = scale[bn]
        w[bn + ".shift"] = shift[bn]
    x = torch.relu(rnd(N, hw, hw, cin))  # NHWC, post-ReLU like a real block input
    g = {}
    out, saved = Bk.bottleneck_fwd(w, x.to(BF), stride, ds)
    gout = rnd(*out.shape)
    from toist_b200 import kernels as K

    gz = K.relu_bwd(gout.to(BF), out)
    gx = Bk.bottleneck_bwd(w, g, {k + ".weight" for k in convs}, gz, saved, stride, ds, True)
    # reference (NCHW fp32) with the *same* bf16-rounded folded weights, so only intermediate rounding differs
    ref = {k: w[k + ".weight"].float().permute(0, 3, 1, 2).contiguous().requires_grad_(True) for k in convs}
    xr = x.permute(0, 3, 1, 2).contiguous().requires_grad_(True)

    def bn(t, name):
        return t + shift[name].view(1, -1, 1, 1)

    def st(t):  # store the activation in bf16 like the kernels do (straight-through gradient)
        return t + (t.to(BF).float() - t).detach()

    y = st(F_.relu(bn(F_.conv2d(xr, ref["conv1"]), "bn1")))
    y = st(F_.relu(bn(F_.conv2d(y, ref["conv2"], stride=stride, padding=1), "bn2")))
    y = bn(F_.conv2d(y, ref["conv3"]), "bn3")
    idt = st(bn(F_.conv2d(xr, ref["downsample.0"], stride=stride), "downsample.1")) if ds else xr
    yr = F_.relu(y + idt)
    yr.backward(gout.permute(0, 3, 1, 2))
    assert rel_err(out.float().permute(0, 3, 1, 2), yr) < FWD_TOL
    # ours returns dL/dx masked by (x > 0) (the ReLU of the producing block is folded in)
    assert rel_err(gx.float().permute(0, 3, 1, 2), xr.grad * (xr > 0)) < TIGHT
    for k in convs:
        # ours: gradient w.r.t. the *unscaled* master weight = scale * d(folded weight)
        want = ref[k].grad * scale[pair[k]][:, None, None, None]
        assert rel_err(g[k + ".weight"], want) < TIGHT, k


def test_input_proj_sequence_layout():
    import torch.nn.functional as F_

    from toist_b200 import blocks as Bk

    torch.manual_seed(4)
    B, h, wd, cin, cout = 3, 7, 9, 2048, 256
    x = rnd(B, h, wd, cin)
    wt, bias = rnd(cout, cin, scale=cin ** -0.5), rnd(cout, scale=0.1)
    out = torch.empty((h * wd * B, cout), dtype=BF, device=DEV)
    Bk.seq_from_nhwc_fwd(x.to(BF), wt.to(BF), bias, out, B)
    xr = x.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr, br = wt.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    yr = F_.conv2d(xr, wr[:, :, None, None], br).flatten(2).permute(2, 0, 1)  # [hw, B, C]
    assert rel_err(out.float().view(h * wd, B, cout), yr) < FWD_TOL
    d = rnd(h * wd * B, cout)
    yr.backward(d.view(h * wd, B, cout))
    g = {}
    dx = Bk.seq_from_nhwc_bwd(g, {"weight", "bias"}, "weight", "bias", d.to(BF), x.to(BF), wt.to(BF), True)
    assert rel_err(dx.float().permute(0, 3, 1, 2), xr.grad) < BWD_TOL
    assert rel_err(g["weight"].view(cout, cin), wr.grad) < BWD_TOL
    assert rel_err(g["bias"], br.grad) < BWD_TOL


def test_heads_linear_layouts():
    import torch.nn.functional as F_

    from toist_b200 import blocks as Bk
    from toist_b200 import kernels as K
    from toist_b200._lib import ACT_SIGMOID

    torch.manual_seed(5)
    L, Q, B, E = 3, 100, 2, 256
    hs = rnd(L, Q * B, E)
    for N, act in ((256, 0), (4, ACT_SIGMOID), (64, 0)):
        wt, bias = rnd(N, E, scale=E ** -0.5), rnd(N, scale=0.1)
        out = Bk.heads_linear_fwd(hs.to(BF), wt.to(BF), bias, L, Q, B, act=act)
        hr = hs.clone().requires_grad_(True)
        wr, br = wt.clone().requires_grad_(True), bias.clone().requires_grad_(True)
        yr = F_.linear(hr.view(L, Q, B, E).transpose(1, 2), wr, br)  # [L, B, Q, N]
        if act:
            yr = yr.sigmoid()
        assert rel_err(out, yr) < 1e-5
        d = rnd(L, B, Q, N)
        yr.backward(d)
        dpre = K.sigmoid_bwd(d.contiguous(), out) if act else d
        d16 = K.cast_pad_bf16(dpre.contiguous(), max(8, N))
        g = {}
        dhs = Bk.heads_linear_bwd(g, {"w", "b"}, "w", "b", d16, hs.to(BF), wt.to(BF), L, Q, B)
        assert rel_err(dhs.float(), hr.grad) < BWD_TOL, N
        assert rel_err(g["w"], wr.grad) < BWD_TOL, N
        assert rel_err(g["b"], br.grad) < BWD_TOL, N


@pytest.mark.parametrize("n,h,w", [(2, 64, 64), (3, 70, 58), (1, 33, 47)])
def test_stem_without_im2col_matches_the_im2col_stem_and_torch(n, h, w):
    """ResNet stem (7x7 stride 2 pad 3 + FrozenBatchNorm + ReLU, torchvision resnet via models/backbone.py:75) as an
    implicit GEMM over the padded NHWC-8 image (kernels.stem_conv7x7: overlapping TMA rows, one tap per kernel row) against
    the round-1 path (explicit im2col matrix + GEMM) and against fp32 torch on the bf16-rounded operands; odd sizes
    exercise the right / bottom border."""
    import torch.nn.functional as F

    from toist_b200 import kernels as K

    g = torch.Generator().manual_seed(7)
    images = torch.randn(n, 3, h, w, generator=g).to(DEV)
    wt = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).to(DEV)
    scale = (torch.rand(64, generator=g) + 0.5).to(DEV)
    shift = torch.randn(64, generator=g).to(DEV)
    ws = (wt * scale.view(-1, 1, 1, 1))
    w7 = torch.zeros(64, 7, 8, 8, dtype=BF, device=DEV)
    w7[:, :, :7, :3] = ws.permute(0, 2, 3, 1).to(BF)
    y = K.stem_conv7x7(images, w7.view(64, 448), shift)
    ho, wo = K.conv_out_size(h, 7, 2, 3), K.conv_out_size(w, 7, 2, 3)
    assert y.shape == (n, ho, wo, 64)
    # im2col path on the same bf16 weights
    sh = torch.zeros(64, 192, dtype=BF, device=DEV)
    sh[:, :147] = ws.permute(0, 2, 3, 1).reshape(64, 147).to(BF)
    y0 = K.linear_fwd(K.stem_im2col(images, 192), sh, shift, act=1).view(n, ho, wo, 64)
    assert rel_err(y.float(), y0.float()) <= 2e-3
    ref = F.relu(F.conv2d(images.to(BF).float(), ws.to(BF).float(), stride=2, padding=3) + shift.view(1, -1, 1, 1))
    assert rel_err(y.float().permute(0, 3, 1, 2), ref) <= 4e-3
