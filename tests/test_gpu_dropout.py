"""Training-mode dropout (reference: nn.Dropout / nn.MultiheadAttention dropout in models/transformer.py:273-283,
337-353,487-492 and the RoBERTa config).  The reference's RNG stream cannot be reproduced, so parity is structural:
the kernels' masks are recovered (the mask is a pure function of (seed, site, flat index)) and a plain torch fp32
implementation of the same layer is run with exactly those masks."""
from __future__ import annotations

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
DEV = "cuda"


def q(t):
    return t.to(BF).float()


def rnd(*shape, scale=1.0):
    return q(torch.randn(*shape, device=DEV) * scale)


def mask_of(shape, drop):
    """keep / (1 - p) factor of every element, recovered by pushing ones through the kernel."""
    from toist_b200 import kernels as K

    return K.dropout(torch.ones(shape, device=DEV), drop)


def test_dropout_kernel_statistics_and_determinism():
    from toist_b200 import kernels as K

    seed = torch.tensor([123456789], dtype=torch.int64, device=DEV)
    x = rnd(4096, 512)
    for p in (0.1, 0.5):
        y = K.dropout(x, (p, seed, 7))
        kept = (y != 0).float().mean().item()
        assert abs(kept - (1 - p)) < 5e-3, kept
        m = y != 0
        assert rel_err(y[m], x[m] / (1 - p)) < 1e-6
        assert torch.equal(y, K.dropout(x, (p, seed, 7)))  # same (seed, site) -> same mask
        assert not torch.equal(y, K.dropout(x, (p, seed, 8)))  # another site -> another mask
        seed2 = seed + 1
        assert not torch.equal(y, K.dropout(x, (p, seed2, 7)))
        xb = x.to(BF)
        yb = K.dropout(xb, (p, seed, 7), res=xb)
        assert torch.equal((yb.float() != xb.float()) | (xb == 0), m | (xb.float() == 0))  # same mask for bf16
    # neighbouring sites and seeds are uncorrelated
    a = mask_of((1 << 20,), (0.5, seed, 1)) > 0
    b = mask_of((1 << 20,), (0.5, seed, 2)) > 0
    assert abs((a & b).float().mean().item() - 0.25) < 5e-3


def test_attention_dropout_forward_backward():
    from toist_b200 import kernels as K

    torch.manual_seed(0)
    sq, sk, b, h, d, p = 100, 233, 2, 8, 32, 0.1
    e = h * d
    seed = torch.tensor([42], dtype=torch.int64, device=DEV)
    drop = (p, seed, 5)
    qq, kk, vv = rnd(sq, b, e).to(BF), rnd(sk, b, e).to(BF), rnd(sk, b, e).to(BF)
    km = torch.zeros(b, sk, dtype=torch.uint8, device=DEV)
    km[1, sk - 7:] = 1
    ctx, probs = K.attention_fwd(qq, kk, vv, km, h, drop=drop, fused=False)  # the unfused path keeps P and dropout(P)
    P, Pd = probs
    ld = P.shape[-1]
    mask = mask_of((b, h, sq, ld), drop)[..., :sk]
    assert abs((mask == 0).float().mean().item() - p) < 5e-3
    # Pd is rounded once from the fp32 probability, P * mask twice: they differ by at most one bf16 ulp
    assert rel_err(Pd[..., :sk].float(), (P[..., :sk].float() * mask).to(BF).float()) < 4e-3

    def heads(t, s):
        return t.float().view(s, b, h, d).permute(1, 2, 0, 3)

    qr, kr, vr = heads(qq, sq).requires_grad_(True), heads(kk, sk).requires_grad_(True), heads(vv, sk).requires_grad_(True)
    att = (qr @ kr.transpose(-1, -2)) * d ** -0.5
    att = att.masked_fill(km.bool()[:, None, None, :], float("-inf")).softmax(-1)
    ref = ((att * mask) @ vr).permute(2, 0, 1, 3).reshape(sq, b, e)
    assert rel_err(ctx.float(), ref) < 6e-3
    dctx = rnd(sq, b, e).to(BF)
    ref.backward(dctx.float())
    dq, dk, dv = torch.empty_like(qq), torch.empty_like(kk), torch.empty_like(vv)
    K.attention_bwd(dctx, qq, kk, vv, probs, h, dq, dk, dv, drop=drop)

    def unheads(t, s):
        return t.permute(2, 0, 1, 3).reshape(s, b, e)

    assert rel_err(dv.float(), unheads(vr.grad, sk)) < 1e-2
    assert rel_err(dq.float(), unheads(qr.grad, sq)) < 2e-2
    assert rel_err(dk.float(), unheads(kr.grad, sk)) < 2e-2


def test_encoder_layer_with_dropout_matches_masked_reference():
    import torch.nn.functional as F

    from test_gpu_blocks import ffn_sd, leaf, ln_sd, mha_sd, split_w
    from toist_b200 import blocks as Bk
    from toist_b200 import kernels as K

    torch.manual_seed(1)
    E, S, B, H, p = 256, 61, 2, 8, 0.1
    dh = E // H
    sd = {**mha_sd("self_attn.", E), **ffn_sd(E, 2048), **ln_sd("norm1.", E), **ln_sd("norm2.", E)}
    x, pos = rnd(S * B, E), rnd(S * B, E)
    km = torch.zeros(B, S, dtype=torch.uint8, device=DEV)
    km[1, S - 5:] = 1
    seed = torch.tensor([2024], dtype=torch.int64, device=DEV)
    drop = Bk.Drop(p, seed, 1000)
    g = {}
    y, saved, _ = Bk.encoder_layer_fwd(split_w(sd), x.to(BF), pos.to(BF), km, H, B, drop)
    dy = rnd(S * B, E)
    dx = K.add_bf16(*Bk.encoder_layer_bwd(split_w(sd), g, set(sd), dy.to(BF), saved, H, B, drop))
    # reference with the recovered masks
    if K.fused_attention_ok(S, S, dh):  # the fused kernel draws its own (paired 16-bit) decisions
        p_att, seed_att, site_att = drop.site(0)
        m_att = K.attention_dropout_mask(B, H, S, S, drop.site(0)).float() / (1.0 - round(p_att * 65536) / 65536.0)
    else:
        ld = (S + 7) // 8 * 8
        m_att = mask_of((B, H, S, ld), drop.site(0))[..., :S]
    m1 = mask_of((S * B, E), drop.site(1))
    mh = mask_of((S * B, 2048), drop.site(2))
    m2 = mask_of((S * B, E), drop.site(3))
    r = leaf(sd)
    xr = x.clone().requires_grad_(True)
    xp = xr + pos
    qk = F.linear(xp, r["self_attn.in_proj_weight"][: 2 * E], r["self_attn.in_proj_bias"][: 2 * E])
    v = F.linear(xr, r["self_attn.in_proj_weight"][2 * E:], r["self_attn.in_proj_bias"][2 * E:])

    def heads(t):
        return t.view(S, B, H, dh).permute(1, 2, 0, 3)

    att = (heads(qk[:, :E]) @ heads(qk[:, E:]).transpose(-1, -2)) * dh ** -0.5
    att = att.masked_fill(km.bool()[:, None, None, :], float("-inf")).softmax(-1) * m_att
    ctx = (att @ heads(v)).permute(2, 0, 1, 3).reshape(S * B, E)
    s1 = xr + F.linear(ctx, r["self_attn.out_proj.weight"], r["self_attn.out_proj.bias"]) * m1
    x1 = F.layer_norm(s1, (E,), r["norm1.weight"], r["norm1.bias"])
    hdn = F.relu(F.linear(x1, r["linear1.weight"], r["linear1.bias"])) * mh
    s2 = x1 + F.linear(hdn, r["linear2.weight"], r["linear2.bias"]) * m2
    yr = F.layer_norm(s2, (E,), r["norm2.weight"], r["norm2.bias"])
    yr.backward(dy)
    assert rel_err(y.float(), yr) < 1e-2
    assert rel_err(dx.float(), xr.grad) < 8e-2
    bad = [(k, rel_err(g[k], r[k].grad)) for k in sd if rel_err(g[k], r[k].grad) > 8e-2]
    assert not bad, bad


def test_model_train_mode_uses_dropout_and_eval_is_deterministic():
    from toist_b200.models import build_model
    from toist_b200.synth import make_args, make_batch, targets_to
    from toist_b200.util.misc import NestedTensor

    torch.manual_seed(0)
    model, criterion, _, wd = build_model(make_args("resnet50", dropout=0.1))
    model.cuda()
    images, mask, captions, targets, pm = make_batch(2, 160, 8, seed=5)
    s = NestedTensor(images.cuda(), mask.cuda())

    def run():
        mc = model(s, captions, encode_and_save=True)
        out = model(s, captions, encode_and_save=False, memory_cache=mc)
        return mc, out

    model.eval()
    with torch.no_grad():
        a, b = run()[1]["pred_logits"].clone(), run()[1]["pred_logits"].clone()
    assert torch.equal(a, b)
    model.train()
    mc, out = run()
    losses = criterion(mc, out, targets_to(targets, "cuda"), pm.cuda(), None)
    total = sum(losses[k] * wd[k] for k in losses if k in wd)
    total.backward()
    assert torch.isfinite(total)
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    with torch.no_grad():
        c = run()[1]["pred_logits"].clone()
        d = run()[1]["pred_logits"].clone()
    assert not torch.equal(c, d), "two training-mode forwards drew the same dropout masks"
    assert rel_err(c, a) > 1e-3, "training-mode forward equals the eval forward: dropout inactive"
