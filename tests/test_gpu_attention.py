"""Fused attention core (csrc/attention.cu: tcgen05 QK^T / PV with the scores in TMEM) against a plain fp32 torch
restatement of what nn.MultiheadAttention computes between in_proj and out_proj (reference models/transformer.py:273,
337-338): softmax(q k^T / sqrt(d) + key_padding_mask) -> dropout -> @ v, forward and backward, on identical bf16 inputs.
Tolerances are norm-wise relative errors; the probabilities are rounded to bf16 before the PV product (as in the
unfused path), so 5e-3 forward / 1.5e-2 backward is the same budget tests/test_gpu_ops.py::test_attention_core uses."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def rel_err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def heads(t, s, b, h, d):
    return t.float().view(s, b, h, d).permute(1, 2, 0, 3)


def unheads(t, s, b, e):
    return t.permute(2, 0, 1, 3).reshape(s, b, e)


def reference(q, k, v, km, h, keep=None, keep_scale=1.0):
    sq, b, e = q.shape
    sk = k.shape[0]
    d = e // h
    qr, kr, vr = (heads(t, s, b, h, d).requires_grad_(True) for t, s in ((q, sq), (k, sk), (v, sk)))
    att = (qr @ kr.transpose(-1, -2)) * d ** -0.5
    if km is not None:
        att = att.masked_fill(km.bool()[:, None, None, :], float("-inf"))
    att = att.softmax(-1)
    if keep is not None:
        att = att * keep.float() * keep_scale
    out = unheads(att @ vr, sq, b, e)
    return out, (qr, kr, vr)


SHAPES = [(416, 416, 2, 8, 32), (100, 416, 2, 8, 32), (100, 100, 2, 8, 32), (16, 16, 3, 12, 64), (233, 233, 2, 8, 32),
          (129, 65, 1, 8, 32), (300, 448, 1, 4, 64)]


@pytest.mark.parametrize("sq,sk,b,h,d", SHAPES)
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_fused_attention_matches_fp32_torch(sq, sk, b, h, d, p_drop):
    from toist_b200 import kernels as K

    torch.manual_seed(sq * 7 + sk)
    e = h * d
    # q and k as column slices of one packed projection (the self-attention layout), v separate
    qk = torch.randn(max(sq, sk), b, 2 * e, device=DEV).to(BF)
    q, k = qk[:sq, :, :e], qk[:sk, :, e:]
    v = torch.randn(sk, b, e, device=DEV).to(BF)
    km = torch.zeros(b, sk, dtype=torch.uint8, device=DEV)
    km[0, sk - 3:] = 1
    if b > 1:
        km[1, ::5] = 1
    drop = keep = None
    ks = 1.0
    if p_drop > 0:
        seed = torch.tensor([1234], dtype=torch.int64, device=DEV)
        drop = (p_drop, seed, 17)
        keep = K.attention_dropout_mask(b, h, sq, sk, drop)
        thr = round(p_drop * 65536)
        ks = 1.0 / (1.0 - thr / 65536.0)
        assert abs(1.0 - keep.float().mean().item() - p_drop) < 1e-2
        other = K.attention_dropout_mask(b, h, sq, sk, (p_drop, seed, 18))
        assert not torch.equal(keep, other)
    assert K.fused_attention_ok(sq, sk, d)
    ctx, saved = K.attention_fwd(q, k, v, km, h, drop=drop, fused=True)
    assert isinstance(saved, K.FusedAttnSaved)
    ref, (qr, kr, vr) = reference(q, k, v, km, h, keep, ks)
    assert torch.isfinite(ctx.float()).all()
    assert rel_err(ctx.float(), ref) < 5e-3
    # row log-sum-exp
    att = (heads(q, sq, b, h, d) @ heads(k, sk, b, h, d).transpose(-1, -2)) * d ** -0.5
    att = att.masked_fill(km.bool()[:, None, None, :], float("-inf"))
    import math

    assert saved.lse.shape == (b, h, sq, 2)  # (raw row maximum of q.k, softmax denominator)
    assert float(saved.lse[..., 1].min()) >= 1.0 - 1e-6 and float(saved.lse[..., 1].max()) <= sk  # e <= 1, e_max = 1
    assert rel_err(saved.lse[..., 0] * d ** -0.5 + saved.lse[..., 1].log(), torch.logsumexp(att, -1)) < 1e-4
    dctx = torch.randn(sq, b, e, device=DEV).to(BF)
    ref.backward(dctx.float())
    dqk = torch.full((max(sq, sk), b, 2 * e), float("nan"), device=DEV, dtype=BF)
    dq, dk = dqk[:sq, :, :e], dqk[:sk, :, e:]
    dv = torch.empty_like(v)
    K.attention_bwd(dctx, q, k, v, saved, h, dq, dk, dv, drop=drop)
    assert rel_err(dv.float(), unheads(vr.grad, sk, b, e)) < 8e-3
    assert rel_err(dq.float(), unheads(qr.grad, sq, b, e)) < 1.5e-2
    assert rel_err(dk.float(), unheads(kr.grad, sk, b, e)) < 1.5e-2
    # masked keys receive exactly zero gradient; nothing is written outside the slices
    assert float(dk[sk - 3:, 0].float().abs().max()) == 0.0 and float(dv[sk - 3:, 0].float().abs().max()) == 0.0
    # deterministic: no atomics anywhere
    dq2, dk2, dv2 = torch.empty_like(dq), torch.empty_like(dk), torch.empty_like(dv)
    K.attention_bwd(dctx, q, k, v, saved, h, dq2, dk2, dv2, drop=drop)
    assert torch.equal(dq2, dq.contiguous()) and torch.equal(dk2, dk.contiguous()) and torch.equal(dv2, dv)


def test_fused_equals_unfused_path():
    from toist_b200 import kernels as K

    torch.manual_seed(3)
    sq, sk, b, h, d = 416, 416, 2, 8, 32
    e = h * d
    q, k, v = (torch.randn(s, b, e, device=DEV).to(BF) for s in (sq, sk, sk))
    km = torch.zeros(b, sk, dtype=torch.uint8, device=DEV)
    km[1, 400:] = 1
    c1, s1 = K.attention_fwd(q, k, v, km, h, fused=True)
    c0, s0 = K.attention_fwd(q, k, v, km, h, fused=False)
    assert rel_err(c1.float(), c0.float()) < 6e-3
    dctx = torch.randn(sq, b, e, device=DEV).to(BF)
    g1 = [torch.empty_like(t) for t in (q, k, v)]
    g0 = [torch.empty_like(t) for t in (q, k, v)]
    K.attention_bwd(dctx, q, k, v, s1, h, *g1)
    K.attention_bwd(dctx, q, k, v, s0, h, *g0)
    for a, b_ in zip(g1, g0):
        assert rel_err(a.float(), b_.float()) < 1.5e-2


def test_unsupported_shapes_use_the_unfused_path():
    from toist_b200 import kernels as K

    assert not K.fused_attention_ok(100, 449, 32)
    assert not K.fused_attention_ok(100, 100, 48)
    torch.manual_seed(0)
    sq, sk, b, h, d = 64, 500, 1, 8, 32
    q, k, v = (torch.randn(s, b, h * d, device=DEV).to(BF) for s in (sq, sk, sk))
    ctx, saved = K.attention_fwd(q, k, v, None, h)
    assert not isinstance(saved, K.FusedAttnSaved)
    ref, _ = reference(q, k, v, None, h)
    assert rel_err(ctx.float(), ref) < 5e-3


def test_fused_attention_backward_is_finite_for_huge_logits():
    """Random-initialised ResNet-101 features reach 1e5, the first encoder layer's logits 1e10 (one ulp = 1e3): the
    softmax is then one-hot, the reference's gradients w.r.t. q and k vanish and w.r.t. v route dO to the winning key.
    The flash-style recompute must not turn rounding differences of such numbers into inf / NaN."""
    from toist_b200 import kernels as K

    torch.manual_seed(5)
    sq, sk, b, h, d = 416, 416, 1, 8, 32
    e = h * d
    q, k = (torch.randn(sq, b, e, device=DEV) * 2.0e4).to(BF), (torch.randn(sk, b, e, device=DEV) * 2.0e4).to(BF)
    v = (torch.randn(sk, b, e, device=DEV) * 1.0e3).to(BF)
    ctx, saved = K.attention_fwd(q, k, v, None, h, fused=True)
    assert torch.isfinite(ctx.float()).all()
    assert float(ctx.float().abs().max()) <= float(v.float().abs().max()) * 1.01  # a convex combination of the values
    dctx = torch.randn(sq, b, e, device=DEV).to(BF)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    K.attention_bwd(dctx, q, k, v, saved, h, dq, dk, dv)
    bad = {n: int((~torch.isfinite(t.float())).sum()) for n, t in (("dq", dq), ("dk", dk), ("dv", dv))}
    assert not any(bad.values()), bad
    # the softmax is one-hot here: dv routes every dO row to its winning key, so it matches fp32 torch closely
    ref, (qr, kr, vr) = reference(q, k, v, None, h)
    ref.backward(dctx.float())
    # (a row whose two best logits are closer than the fp32 rounding of a 1e10 dot product may pick the other key)
    assert rel_err(ctx.float(), ref.detach()) < 5e-2
    assert rel_err(dv.float(), unheads(vr.grad, sk, b, e)) < 5e-2
