"""GPU parity of the small (HBM / latency bound) kernels against plain torch fp32 math and the oracle.

Tolerances: fp32 kernels 1e-5 norm-wise; kernels whose *output* is bf16 are compared after rounding the fp32
reference to bf16 (<= 1 bf16 ulp = 2**-8 relative per element, 3e-3 norm-wise).  Integer outputs are bit exact.
"""
from __future__ import annotations

import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import max_err, rel_err

pytestmark = pytest.mark.gpu

BF = torch.bfloat16
DEV = "cuda"


@pytest.fixture(scope="module")
def K():
    from toist_b200 import kernels

    return kernels


def rn(*shape, scale=1.0, dtype=torch.float32, seed=None):
    if seed is not None:
        torch.manual_seed(seed)
    return (torch.randn(*shape, device=DEV) * scale).to(dtype)


# ------------------------------------------------------------------------------------------------ elementwise
def test_casts_and_add(K):
    torch.manual_seed(0)
    for n in (8, 1000, 2048 * 3 + 5):
        x = rn(n)
        assert torch.equal(K.cast_bf16(x), x.to(BF))
        assert torch.equal(K.cast_f32(x.to(BF)), x.to(BF).float())
        a, b, c = rn(n, dtype=BF), rn(n, dtype=BF), rn(n, dtype=BF)
        assert torch.equal(K.add_bf16(a, b), (a.float() + b.float()).to(BF))
        assert torch.equal(K.add_bf16(a, b, c), (a.float() + b.float() + c.float()).to(BF))


def test_weight_prep(K):
    torch.manual_seed(1)
    prep = K.WeightPrep(torch.device(DEV))
    srcs = [rn(64, 147), rn(256, 64), rn(768, 256), rn(5, 7)]
    scales = [torch.rand(64, device=DEV) + 0.5, None, None, torch.rand(5, device=DEV)]
    ldds = [192, 64, 256, 8]
    dsts = [torch.zeros(s.shape[0], l, device=DEV, dtype=BF) for s, l in zip(srcs, ldds)]
    for s, d, sc, l in zip(srcs, dsts, scales, ldds):
        prep.add(s, d, s.shape[0], s.shape[1], l, sc)
    prep.run()
    for s, d, sc in zip(srcs, dsts, scales):
        ref = s if sc is None else s * sc[:, None]
        assert torch.equal(d[:, : s.shape[1]], ref.to(BF))
        assert bool((d[:, s.shape[1]:] == 0).all())


def test_stem_im2col_and_maxpool(K):
    torch.manual_seed(2)
    img = rn(2, 3, 38, 50)
    p = K.stem_im2col(img, 192)
    ref = F.unfold(img, kernel_size=7, padding=3, stride=2)  # [N, 3*49, L] with column order (c, ky, kx)
    n, _, L = ref.shape
    ref = ref.view(n, 3, 49, L).permute(0, 3, 2, 1).reshape(n * L, 147)  # -> (ky*7+kx)*3 + c
    assert torch.equal(p[:, :147], ref.to(BF))
    assert bool((p[:, 147:] == 0).all())
    x = rn(2, 19, 23, 64, dtype=BF)
    y = K.maxpool3x3s2(x)
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(y.float(), ref)


def test_reductions_and_broadcasts(K):
    torch.manual_seed(3)
    for dt in (torch.float32, BF):
        x = rn(1111, 200, dtype=dt)
        out = torch.zeros(200, device=DEV)
        K.colsum(x, out)
        assert rel_err(out, x.float().sum(0)) < 1e-5
        x3 = rn(7, 5, 96, dtype=dt)
        assert rel_err(K.sum_mid(x3), x3.float().sum(1)) < 1e-6
    q = rn(100, 256)
    assert torch.equal(K.bcast_mid(q, 3), q.to(BF)[:, None, :].expand(100, 3, 256))
    x = rn(2, 24, 5, 7)
    nhwc = K.nchw_to_nhwc(x)
    assert torch.equal(nhwc, x.permute(0, 2, 3, 1).to(BF))
    assert torch.equal(K.nhwc_to_nchw(nhwc), x.to(BF).float())


def test_activation_grads(K):
    torch.manual_seed(4)
    n = 5000
    dy, pre = rn(n, dtype=BF), rn(n, dtype=BF, scale=2.0)
    p = pre.float().requires_grad_(True)
    F.gelu(p).backward(dy.float())
    assert rel_err(K.gelu_bwd(dy, pre), p.grad) < 3e-3
    y, dy2 = rn(n, dtype=BF), rn(n, dtype=BF)
    assert torch.equal(K.relu_bwd(dy, y), (dy.float() * (y.float() > 0)).to(BF))
    assert torch.equal(K.relu_bwd(dy, y, dy2), ((dy.float() + dy2.float()) * (y.float() > 0)).to(BF))
    s = torch.rand(n, device=DEV)
    g = rn(n)
    assert rel_err(K.sigmoid_bwd(g, s), g * s * (1 - s)) < 1e-6


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("n,rows,dt", [(256, 3328, BF), (768, 128, torch.float32), (64, 37, torch.float32)])
def test_layernorm(K, n, rows, dt):
    torch.manual_seed(5)
    x = rn(rows, n, dtype=dt, scale=3.0)
    g, b = rn(n) * 0.2 + 1, rn(n) * 0.1
    y16, y32, mean, rstd = K.layernorm_fwd(x, g, b, 1e-5, want_f32=True)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (n,), gr, br, 1e-5)
    assert rel_err(y32, ref) < 1e-5
    assert torch.equal(y16, y32.to(BF))
    dy = rn(rows, n, dtype=BF)
    dy2 = rn(rows, n, dtype=BF)
    ref.backward(dy.float() + dy2.float())
    dg, db = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    dx = K.layernorm_bwd(dy, x, mean, rstd, g, dy2=dy2, dgamma=dg, dbeta=db, dx_dtype=torch.float32)
    assert rel_err(dx, xr.grad) < 1e-4
    assert rel_err(dg, gr.grad) < 1e-4
    assert rel_err(db, br.grad) < 1e-4


def test_l2norm(K):
    torch.manual_seed(6)
    x = rn(4800, 64)
    y, nrm = K.l2norm_fwd(x)
    xr = x.clone().requires_grad_(True)
    ref = F.normalize(xr, p=2, dim=-1)
    assert rel_err(y, ref) < 1e-6
    dy = rn(4800, 64)
    ref.backward(dy)
    assert rel_err(K.l2norm_bwd(dy, y, nrm), xr.grad) < 1e-5


def test_attn_softmax(K):
    torch.manual_seed(7)
    b, h, sq, sk = 2, 8, 100, 233
    ld = (sk + 7) // 8 * 8
    s = rn(b, h, sq, ld, scale=3.0)
    km = torch.zeros(b, sk, dtype=torch.uint8, device=DEV)
    km[0, 200:] = 1
    km[1, 5:9] = 1
    p = torch.empty(b, h, sq, ld, dtype=BF, device=DEV)
    from toist_b200 import _lib

    L = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.toist_attn_softmax_fwd(s.data_ptr(), km.data_ptr(), p.data_ptr(), None, b * h * sq, sk, ld, ld, h * sq,
                                        0.0, None, 0, st))
    ref = s[..., :sk].masked_fill(km.bool()[:, None, None, :], float("-inf")).softmax(-1)
    assert rel_err(p[..., :sk].float(), ref) < 3e-3
    assert bool((p[..., sk:] == 0).all())
    dp = rn(b, h, sq, ld)
    ds = torch.empty_like(p)
    _lib.check(L.toist_attn_softmax_bwd(dp.data_ptr(), p.data_ptr(), ds.data_ptr(), b * h * sq, sk, ld, ld, 0.25, 0.0,
                                        None, 0, st))
    pf = p[..., :sk].float()
    refd = pf * (dp[..., :sk] - (dp[..., :sk] * pf).sum(-1, keepdim=True)) * 0.25
    assert rel_err(ds[..., :sk].float(), refd) < 4e-3


def test_pos_sine(K):
    from oracle import model as O

    torch.manual_seed(8)
    mask = torch.zeros(3, 15, 20, dtype=torch.bool)
    mask[1, :, 16:] = True
    mask[2, 11:, :] = True
    p32, p16 = K.pos_sine(mask.to(DEV).to(torch.uint8), 128)
    ref = O.position_sine(mask, 128).flatten(2).permute(2, 0, 1)  # [HW, B, 256]
    assert max_err(p32, ref) < 2e-5
    assert max_err(p16.float(), ref) < 5e-3


def test_roberta_embeddings(K):
    torch.manual_seed(9)
    V, P, E = 500, 40, 768
    word, pos, typ = rn(V, E), rn(P, E), rn(1, E)
    ids = torch.randint(3, V, (4, 16), device=DEV)
    ids[1, 10:] = 1
    ids[3, 5:] = 1
    out, pos_ids = K.embed_gather(ids, word, pos, typ)
    nonpad = ids.ne(1).int()
    pid = (torch.cumsum(nonpad, 1) * nonpad).long() + 1
    ref = word[ids] + pos[pid] + typ[0]
    assert torch.equal(pos_ids.view(4, 16).long(), pid)
    assert rel_err(out, ref.view(-1, E)) < 1e-6
    dx = rn(64, E)
    dw, dp, dt = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros(E, device=DEV)
    K.embed_scatter(dx, ids, pos_ids, dw, dp, dt)
    rw = torch.zeros_like(word).index_add_(0, ids.flatten(), dx)
    rp = torch.zeros_like(pos).index_add_(0, pid.flatten(), dx)
    assert rel_err(dw, rw) < 1e-5 and rel_err(dp, rp) < 1e-5 and rel_err(dt, dx.sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------ attention (composite)
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("sq,sk,b,h,d", [(416, 416, 2, 8, 32), (100, 233, 2, 8, 32), (16, 16, 3, 12, 64)])
def test_attention_core(K, sq, sk, b, h, d, fused):
    torch.manual_seed(10)
    e = h * d
    q, k, v = rn(sq, b, e, dtype=BF), rn(sk, b, e, dtype=BF), rn(sk, b, e, dtype=BF)
    km = torch.zeros(b, sk, dtype=torch.uint8, device=DEV)
    km[0, sk - 3:] = 1
    ctx, probs = K.attention_fwd(q, k, v, km, h, fused=fused)

    def heads(t, s):
        return t.float().view(s, b, h, d).permute(1, 2, 0, 3)

    qr, kr, vr = heads(q, sq).requires_grad_(True), heads(k, sk).requires_grad_(True), heads(v, sk).requires_grad_(True)
    att = (qr @ kr.transpose(-1, -2)) * d ** -0.5
    att = att.masked_fill(km.bool()[:, None, None, :], float("-inf")).softmax(-1)
    ref = (att @ vr).permute(2, 0, 1, 3).reshape(sq, b, e)
    assert rel_err(ctx.float(), ref) < 5e-3
    dctx = rn(sq, b, e, dtype=BF)
    ref.backward(dctx.float())
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    K.attention_bwd(dctx, q, k, v, probs, h, dq, dk, dv)

    def unheads(t, s):
        return t.permute(2, 0, 1, 3).reshape(s, b, e)

    assert rel_err(dv.float(), unheads(vr.grad, sk)) < 8e-3
    assert rel_err(dq.float(), unheads(qr.grad, sq)) < 1.5e-2
    assert rel_err(dk.float(), unheads(kr.grad, sk)) < 1.5e-2


# ------------------------------------------------------------------------------------------------ matcher
def _random_problem(L, B, Q, C, counts, seed):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(L, B, Q, C, generator=g) * 2
    boxes = torch.cat([torch.rand(L, B, Q, 2, generator=g) * 0.5 + 0.25, torch.rand(L, B, Q, 2, generator=g) * 0.3 + 0.05], -1)
    targets = []
    for n in counts:
        tb = torch.cat([torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.3 + 0.05], -1)
        targets.append({"boxes": tb, "labels": torch.ones(n, dtype=torch.long)})
    T = sum(counts)
    pm = torch.zeros(T, C)
    for r in range(T):
        lo = int(torch.randint(1, 8, (1,), generator=g))
        hi = lo + int(torch.randint(1, 6, (1,), generator=g))
        pm[r, lo:hi] = 1
    pm = pm / (pm.sum(-1, keepdim=True) + 1e-6)
    return logits, boxes, targets, pm


def _pad_targets(targets, pm, tmax):
    B = len(targets)
    C = pm.shape[1]
    tb = torch.zeros(B, tmax, 4)
    pp = torch.zeros(B, tmax, C)
    cnt = torch.zeros(B, dtype=torch.int32)
    off = 0
    for i, t in enumerate(targets):
        n = len(t["boxes"])
        tb[i, :n] = t["boxes"]
        pp[i, :n] = pm[off:off + n]
        cnt[i] = n
        off += n
    return tb.to(DEV), cnt.to(DEV), pp.to(DEV)


@pytest.mark.parametrize("seed", range(6))
def test_matcher_cost_and_device_lsap(K, seed):
    from oracle import model as O

    L, B, Q, C = 3, 4, 100, 256
    counts = [[1, 2, 3, 4], [0, 5, 1, 9], [4, 4, 0, 0], [7, 1, 2, 3], [1, 1, 1, 1], [12, 3, 0, 6]][seed]
    tmax = max(max(counts), 1)
    logits, boxes, targets, pm = _random_problem(L, B, Q, C, counts, seed)
    tb, cnt, pp = _pad_targets(targets, pm, tmax)
    cost = K.match_cost(logits.to(DEV), boxes.to(DEV), tb, cnt, pp, 1.0, 5.0, 2.0)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    mq = K.lsap_device(cost, cnt, flags).cpu()
    assert int(flags.item()) == 0
    tgt_all = torch.cat([t["boxes"] for t in targets]) if sum(counts) else torch.zeros(0, 4)
    for l in range(L):
        ref_c = O.matcher_cost(logits[l], boxes[l], tgt_all, pm) if sum(counts) else None
        ref_idx = O.hungarian_match(logits[l], boxes[l], targets, pm)
        off = 0
        for b, n in enumerate(counts):
            if n:
                blk = ref_c[b][:, off:off + n]
                assert max_err(cost[l, b, :, :n], blk) < 2e-6
            off += n
            rows, cols = ref_idx[b]
            got = mq[l, b, :n]
            # reference: rows ascending, cols[k] = target of query rows[k]
            want = torch.full((n,), -1, dtype=torch.int32)
            want[cols] = rows.to(torch.int32)
            assert torch.equal(got, want), f"layer {l} image {b}: {got.tolist()} vs {want.tolist()}"
            assert bool((mq[l, b, n:] == -1).all())


def test_lsap_host_matches_scipy(K):
    from scipy.optimize import linear_sum_assignment

    rng = np.random.RandomState(0)
    shapes = [(100, 4), (4, 100), (97, 97), (1, 1), (5, 0), (0, 5), (30, 1024), (100, 1), (3, 3)]
    for nr, nc in shapes:
        for trial in range(3):
            c = rng.rand(nr, nc)
            if trial == 1 and c.size:
                c = np.round(c * 4) / 4  # heavy ties
            if trial == 2 and c.size:
                c = c.astype(np.float32).astype(np.float64)
            r0, c0 = linear_sum_assignment(c)
            r1, c1 = K.lsap_host(c)
            assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist(), (nr, nc, trial)
    with pytest.raises(ValueError):
        K.lsap_host(np.array([[np.nan, 1.0], [1.0, 2.0]]))
    with pytest.raises(ValueError):
        K.lsap_host(np.array([[np.inf, np.inf], [np.inf, np.inf]]))


def test_device_lsap_flags_nan(K):
    cost = torch.zeros(1, 1, 10, 2, device=DEV)
    cost[0, 0, 3, 1] = float("nan")
    cnt = torch.tensor([2], dtype=torch.int32, device=DEV)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    K.lsap_device(cost, cnt, flags)
    assert int(flags.item()) == 1


# ------------------------------------------------------------------------------------------------ criterion kernels
def test_criterion_kernels_against_oracle(K):
    from oracle import model as O
    from toist_b200 import _lib

    L, B, Q, C, Kt, D = 3, 4, 100, 256, 16, 64
    counts = [1, 2, 3, 4]
    tmax = 4
    logits, boxes, targets, pm = _random_problem(L, B, Q, C, counts, 123)
    g = torch.Generator().manual_seed(5)
    pq = F.normalize(torch.randn(L, B, Q, D, generator=g), dim=-1)
    pt = F.normalize(torch.randn(B, Kt, D, generator=g), dim=-1)
    tb, cnt, pp = _pad_targets(targets, pm, tmax)
    nb = float(sum(counts))
    # token spans: target t of image b is positive on tokens [1 + t, 3 + t]
    tok_pos = torch.zeros(B, tmax, Kt, dtype=torch.uint8)
    for b, n in enumerate(counts):
        for t in range(n):
            tok_pos[b, t, 1 + t: 4 + t] = 1

    class Tok:
        def char_to_token(self, i, c=None):
            return (i if c is None else c) + 1

    for b, n in enumerate(counts):
        targets[b]["tokens_positive"] = [[[t, t + 3]] for t in range(n)]  # chars [t, t+3) -> tokens t+1 .. t+3

    lg = logits.to(DEV).requires_grad_(True)
    bx = boxes.to(DEV).requires_grad_(True)
    pqd = pq.to(DEV).requires_grad_(True)
    ptd = pt.to(DEV).requires_grad_(True)
    cost = K.match_cost(lg.detach(), bx.detach(), tb, cnt, pp, 1.0, 5.0, 2.0)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    mq = K.lsap_device(cost, cnt, flags)
    nbt = torch.tensor([nb], device=DEV)
    Lb = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    row_loss = torch.empty(L, B, Q, device=DEV)
    dlogits = torch.empty(L, B, Q, C, device=DEV)
    _lib.check(Lb.toist_token_ce(lg.data_ptr(), mq.data_ptr(), cnt.data_ptr(), pp.data_ptr(), nbt.data_ptr(),
                                 row_loss.data_ptr(), dlogits.data_ptr(), L, B, Q, C, tmax, 0.1, st))
    pl1, pgi = torch.empty(L, B, tmax, device=DEV), torch.empty(L, B, tmax, device=DEV)
    db1, db2 = torch.empty(L, B, Q, 4, device=DEV), torch.empty(L, B, Q, 4, device=DEV)
    _lib.check(Lb.toist_box_loss(bx.data_ptr(), mq.data_ptr(), cnt.data_ptr(), tb.data_ptr(), nbt.data_ptr(),
                                 pl1.data_ptr(), pgi.data_ptr(), db1.data_ptr(), db2.data_ptr(), L, B, Q, tmax, st))
    card = torch.empty(L, B, dtype=torch.int32, device=DEV)
    _lib.check(Lb.toist_cardinality(lg.data_ptr(), card.data_ptr(), L, B, Q, C, st))
    img_loss = torch.empty(L, B, device=DEV)
    dpq, dpt = torch.empty(L, B, Q, D, device=DEV), torch.empty(L, B, Kt, D, device=DEV)
    tpd = tok_pos.to(DEV)
    _lib.check(Lb.toist_contrastive_align(pqd.data_ptr(), ptd.data_ptr(), mq.data_ptr(), cnt.data_ptr(), tpd.data_ptr(),
                                          nbt.data_ptr(), img_loss.data_ptr(), dpq.data_ptr(), dpt.data_ptr(), L, B, Q,
                                          Kt, D, tmax, 0.07, st))
    out = torch.empty(5, L, device=DEV)
    _lib.check(Lb.toist_criterion_reduce(row_loss.data_ptr(), pl1.data_ptr(), pgi.data_ptr(), card.data_ptr(),
                                         img_loss.data_ptr(), cnt.data_ptr(), nbt.data_ptr(), flags.data_ptr(), out.data_ptr(), L, B, Q,
                                         tmax, st))
    out = out.cpu()
    for l in range(L):
        lgl, bxl = lg[l].cpu(), bx[l].cpu()
        idx = O.hungarian_match(lgl.detach(), bxl.detach(), targets, pm)
        lg_r = lgl.detach().clone().requires_grad_(True)
        bx_r = bxl.detach().clone().requires_grad_(True)
        pq_r = pq[l].clone().requires_grad_(True)
        pt_r = pt.clone().requires_grad_(True)
        ce = O.loss_labels(lg_r, targets, pm, idx, nb, 0.1)
        l1, gi = O.loss_boxes(bx_r, targets, idx, nb)
        ca = O.loss_contrastive_align(pq_r, pt_r, Tok(), targets, idx, nb, 0.07)
        cd = O.loss_cardinality(lg_r, targets)
        assert abs(out[0, l].item() - ce.item()) < 1e-4 * max(1, abs(ce.item()))
        assert abs(out[1, l].item() - l1.item()) < 1e-5 * max(1, abs(l1.item()))
        assert abs(out[2, l].item() - gi.item()) < 1e-5 * max(1, abs(gi.item()))
        assert abs(out[3, l].item() - cd.item()) < 1e-6
        assert abs(out[4, l].item() - ca.item()) < 1e-4 * max(1, abs(ca.item()))
        ce.backward()
        assert rel_err(dlogits[l], lg_r.grad) < 1e-4
        l1.backward(retain_graph=True)
        assert rel_err(db1[l], bx_r.grad) < 1e-5
        bx_r.grad = None
        gi.backward()
        assert rel_err(db2[l], bx_r.grad) < 1e-4
        ca.backward()
        assert rel_err(dpq[l], pq_r.grad) < 1e-4
        assert rel_err(dpt[l], pt_r.grad) < 1e-4


def test_scale_layers(K):
    from toist_b200 import _lib

    x = rn(6, 1000)
    g = rn(6)
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.load().toist_scale_layers(x.data_ptr(), g.data_ptr(), y.data_ptr(), 6, 1000, 0, st))
    assert rel_err(y, x * g[:, None]) < 1e-6
    z = torch.empty(1000, device=DEV)
    _lib.check(_lib.load().toist_scale_layers(x.data_ptr(), g.data_ptr(), z.data_ptr(), 6, 1000, 1, st))
    assert rel_err(z, (x * g[:, None]).sum(0)) < 1e-5


def test_embed_rows_merge_is_the_sparse_form_of_the_dense_mean():
    """Receive side of the sparse word-embedding gradient exchange (util/dist.FlatGradSync._sparse_exchange; replaces
    the dense all-reduce of main.py:336 for that table): table[id] = scale * sum of the gathered rows with that id in
    list order; padding-id slots are skipped, untouched rows keep what they held, the result is the same on every call
    (fixed summation order: bit-identical across ranks)."""
    from toist_b200 import kernels as K

    g = torch.Generator().manual_seed(3)
    vocab, e, world, slots, pad = 50265, 768, 4, 64, 1
    ids = torch.full((world * slots,), pad, dtype=torch.int64)
    rows = torch.randn(world * slots, e, generator=g)
    for r in range(world):  # every rank fills a different number of its fixed slots; ids repeat within and across ranks
        n = 10 + 7 * r
        ids[r * slots: r * slots + n] = torch.randint(2, 40, (n,), generator=g)
    table = torch.full((vocab, e), 7.0)
    want = table.clone()
    used = ids != pad
    acc = torch.zeros(vocab, e, dtype=torch.float64)
    acc.index_add_(0, ids[used], rows[used].double())
    touched = torch.zeros(vocab, dtype=torch.bool)
    touched[ids[used]] = True
    want[touched] = (acc[touched] / world).float()
    outs = []
    for _ in range(2):
        t = table.clone().cuda()
        K.embed_rows_merge(t, ids.cuda(), rows.cuda(), pad, 1.0 / world)
        outs.append(t.cpu())
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[0][~touched], table[~touched])
    assert max_err(outs[0][touched], want[touched]) <= 1e-5
