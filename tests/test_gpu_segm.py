"""DETRsegm mask branch + mask losses on the GPU (BASELINE config 3) against the golden
fixture frozen from the unmodified reference and against the fp32 oracle.

Tolerances: the detector runs in bf16, the mask head stores NHWC bf16 maps between its five
conv + GroupNorm stages; pred_masks are compared norm-wise at the bf16 budget, the mask losses (fp32 kernels) relative
to the reference value, and the mask-branch gradients norm-wise."""
from __future__ import annotations

from pathlib import Path

import pytest
import torch

from conftest import rel_err
from toist_b200.synth import make_args, make_batch, targets_to

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _build(frozen=True, **over):
    from toist_b200.models import build_model

    torch.manual_seed(0)
    args = make_args("resnet50", masks=True, mask_model="smallconv", frozen_weights="unused" if frozen else None,
                     aux_loss=False, contrastive_align_loss=False, **over)
    model, criterion, _, wd = build_model(args)
    return model.cuda().eval(), criterion, wd


def _step(model, criterion, wd, batch, grad=True):
    from toist_b200.util.misc import NestedTensor

    images, mask, captions, targets, pm = batch
    s = NestedTensor(images.cuda(), mask.cuda())
    model.zero_grad(set_to_none=True)
    with torch.enable_grad() if grad else torch.no_grad():
        mc = model(s, captions, encode_and_save=True)
        out = model(s, captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, targets_to(targets, "cuda"), pm.cuda(), None)
        if grad:
            sum(losses[k] * wd[k] for k in losses if k in wd).backward()
    return mc, out, losses


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD / "config3_r50_segm_small.pt", weights_only=False)


def test_mask_branch_matches_reference_golden(gold):
    model, criterion, wd = _build(frozen=True)
    b = gold["batch"]
    batch = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"], masks=True)
    mc, out, losses = _step(model, criterion, wd, batch)
    assert out["pred_masks"].shape == gold["pred_masks"].shape and out["pred_masks"].dtype == torch.float32
    assert rel_err(out["pred_masks"], gold["pred_masks"]) < 3e-2
    assert set(losses) == set(gold["losses"])
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    # frozen detector: only the mask branch trains (models/segmentation.py:28-31)
    assert sorted(grads) == sorted(gold["grads"])
    idx = criterion.last_indices()[-1]
    same = all(r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist() for (r0, c0), (r1, c1) in zip(idx, gold["indices"]))
    if same:
        ref_losses, ref_grads = gold["losses"], gold["grads"]
    else:
        # random-init costs are near ties and the detector runs in bf16: when the assignment of some image differs
        # from the reference's, differentiate the oracle (pinned to this golden on CPU) under OUR assignment instead
        ref_losses, ref_grads = _oracle_losses_and_grads(model, batch, wd, idx)
    for k in ("loss_mask", "loss_dice", "loss_bbox", "loss_giou", "loss_ce"):
        v = float(ref_losses[k])
        assert abs(float(losses[k]) - v) <= 2e-2 * max(1.0, abs(v)), (k, float(losses[k]), v)
    worst = []
    for k, g in ref_grads.items():
        if k == "bbox_attention.k_linear.bias":  # mathematically zero (softmax shift invariance)
            assert float(grads[k].abs().max()) <= 5e-2 * float(grads["bbox_attention.q_linear.bias"].abs().max() + 1e-30)
            continue
        worst.append((rel_err(grads[k], g), k))
    worst.sort(reverse=True)
    # against pure fp32: five conv + GroupNorm + ReLU stages deep, ReLU-mask flips dominate (tight check on identical
    # inputs: test_mask_stage_against_fp32_torch)
    assert worst[0][0] < 0.15, worst[:6]
    assert worst[len(worst) // 2][0] < 8e-2, worst[len(worst) // 2]


def _oracle_losses_and_grads(model, batch, wd, idx):
    from oracle import model as O

    images, mask, captions, targets, pm = batch
    sd = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    for k in trainable:
        sd[k].requires_grad_(True)
    tokd = model.detr.transformer.tokenizer(captions)
    cfg = O.Config(backbone="resnet50", prefix="detr.", aux_loss=False, contrastive_align_loss=False)
    mc = O.encode(sd, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
    out = O.decode(sd, cfg, mc)
    out["pred_masks"] = O.decode_masks(sd, cfg, mc, out)
    losses, _ = O.criterion(cfg, out, tokd, targets, pm, masks=True, forced_indices=[idx])
    sum(losses[k] * wd[k] for k in losses if k in wd).backward()
    return {k: float(v) for k, v in losses.items()}, {k: sd[k].grad for k in trainable if sd[k].grad is not None}


def test_mask_losses_on_reference_predictions(gold):
    """loss_masks in isolation: the reference's own fp32 pred_masks and assignment through our fused
    upsample + focal + dice kernel."""
    from toist_b200 import kernels as K

    b = gold["batch"]
    _, _, _, targets, _ = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"], masks=True)
    from toist_b200.models.mdetr import pack_target_masks

    B = len(targets)
    t_max = max(len(t["boxes"]) for t in targets)
    tm = pack_target_masks(targets, t_max, "cuda")
    mq = torch.full((B, t_max), -1, dtype=torch.int32)
    for i, (src, tgt) in enumerate(gold["indices"]):
        mq[i, tgt] = src.int()
    count = torch.tensor([len(t["boxes"]) for t in targets], dtype=torch.int32, device="cuda")
    nb = torch.tensor([float(sum(len(t["boxes"]) for t in targets))], device="cuda")
    pred = gold["pred_masks"].cuda()
    out, sums = K.mask_loss_fwd(pred, tm, mq.cuda(), count, nb)
    assert abs(float(out[0]) - gold["losses"]["loss_mask"]) < 1e-5 * max(1.0, gold["losses"]["loss_mask"])
    assert abs(float(out[1]) - gold["losses"]["loss_dice"]) < 1e-5
    # gradient of loss_mask + loss_dice w.r.t. pred_masks against autograd through the oracle's restatement
    from oracle import model as O

    p = gold["pred_masks"].clone().requires_grad_(True)
    lm, ld = O.loss_masks(p, targets, gold["indices"], float(nb))
    (lm + ld).backward()
    d = K.mask_loss_bwd(pred, tm, mq.cuda(), count, sums, nb, torch.ones(2, device="cuda"))
    assert rel_err(d, p.grad) < 1e-4


def test_unfrozen_detector_receives_gradients_through_the_mask_branch():
    """Without --frozen_weights the mask losses back-propagate into hs, the encoder memory, src_proj and the three
    backbone maps; compare every gradient with the fp32 oracle."""
    from oracle import model as O

    model, criterion, wd = _build(frozen=False)
    batch = make_batch(2, 128, 8, seed=11, pad=True, masks=True)
    sd_cpu = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    mc, out, losses = _step(model, criterion, wd, batch)
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    idx = criterion.last_indices()[-1]

    images, mask, captions, targets, pm = batch
    tokd = model.detr.transformer.tokenizer(captions)
    cfg = O.Config(backbone="resnet50", prefix="detr.", aux_loss=False, contrastive_align_loss=False)
    for k in trainable:
        sd_cpu[k].requires_grad_(True)
    omc = O.encode(sd_cpu, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
    oout = O.decode(sd_cpu, cfg, omc)
    oout["pred_masks"] = O.decode_masks(sd_cpu, cfg, omc, oout)
    ol, _ = O.criterion(cfg, oout, tokd, targets, pm, masks=True, forced_indices=[idx])
    sum(ol[k] * wd[k] for k in ol if k in wd).backward()
    assert rel_err(out["pred_masks"], oout["pred_masks"]) < 3e-2
    for k in ("loss_mask", "loss_dice"):
        assert abs(float(losses[k]) - float(ol[k])) < 2e-2 * max(1.0, abs(float(ol[k])))
    errs = []
    for k in trainable:
        og = sd_cpu[k].grad
        if og is None:
            continue
        assert k in grads, f"missing gradient {k}"
        if float(og.abs().max()) < 1e-12 or "key.bias" in k or k.endswith("k_linear.bias"):
            continue
        errs.append((rel_err(grads[k], og), k))
    assert len(errs) > 250
    errs.sort()
    print('median', errs[len(errs) // 2], 'worst', errs[-8:])
    # the detector gradients sit at the bf16 floor quantified in test_gpu_model.py (ReLU-mask flips at random init
    # put tens of percent on deep backbone tensors for any bf16 implementation); the new paths are checked tightly
    # on identical inputs in test_mask_stage_against_fp32_torch below
    assert errs[len(errs) // 2][0] < 0.15, errs[len(errs) // 2]
    mask_branch = [e for e in errs if e[1].startswith(("bbox_attention.", "mask_head."))]
    assert max(mask_branch)[0] < 0.15, sorted(mask_branch)[-5:]


def test_mask_branch_cuda_graph_replay_matches_eager():
    model, criterion, wd = _build(frozen=True)
    batches = [make_batch(2, 128, 8, seed=s, pad=True, masks=True) for s in (3, 4)]

    def run(b):
        _, out, losses = _step(model, criterion, wd, b)
        return out["pred_masks"].clone(), float(losses["loss_mask"]), float(losses["loss_dice"]), \
            {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    eager = [run(b) for b in batches]
    model.enable_cuda_graphs(True)
    criterion.enable_cuda_graphs(True)
    for _ in range(2):
        for (pm0, lm0, ld0, g0), b in zip(eager, batches):
            pm1, lm1, ld1, g1 = run(b)
            assert torch.equal(pm0, pm1), float((pm0 - pm1).abs().max())
            assert abs(lm0 - lm1) <= 1e-6 * max(1.0, abs(lm0)) and abs(ld0 - ld1) <= 1e-6
            for k in g0:
                assert bool(torch.isfinite(g0[k]).all()) and bool(torch.isfinite(g1[k]).all()), \
                    (k, int((~torch.isfinite(g0[k])).sum()), int((~torch.isfinite(g1[k])).sum()), g0[k].numel())
                # mask_loss_bwd scatters through fp32 atomics (order varies run to run) and the maps in between are
                # bf16: replay and eager agree norm-wise, not bit for bit
                # bbox_attention gradients pass the softmax backward P * (dP - sum P dP): a cancelling difference of
                # atomically accumulated, bf16-stored values whose result is ~1e-8 here; its run-to-run noise was
                # measured at 3e-3 .. 7e-3 (the test passed or failed with the atomics' arrival order)
                tol = 2e-2 if "bbox_attention." in k else 5e-3
                if k == "bbox_attention.k_linear.bias":
                    # exact gradient is zero (softmax shift invariance): what is left is rounding noise, ~1e-10 against
                    # 1e-8 .. 1e-7 for q_linear.bias; only its smallness is meaningful
                    assert float(g1[k].abs().max()) <= 5e-2 * float(g1["bbox_attention.q_linear.bias"].abs().max() + 1e-30)
                    continue
                assert rel_err(g1[k], g0[k]) < tol or float(g0[k].abs().max()) < 1e-10, (k, rel_err(g1[k], g0[k]))


def test_mask_stage_against_fp32_torch():
    """The mask branch as one stage (toist_b200/maskhead.py) on identical bf16-representable inputs and weights against
    the oracle's fp32 restatement (oracle.attention_map + oracle.mask_head) run with torch on the GPU: forward, every
    parameter gradient and every input gradient (hs, encoder memory, src_proj, the three FPN maps)."""
    from types import SimpleNamespace

    from oracle import model as O
    from toist_b200 import maskhead as MH

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(9)
    BF = torch.bfloat16
    dev = "cuda"
    B, Q, E, NH, L, h, w, n_text = 2, 7, 256, 8, 2, 5, 6, 4
    hw, S = h * w, h * w + n_text

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev) * scale).to(BF).float()

    dims = [E + NH, 128, 64, 32, 16, 4]
    sd = {"bbox_attention.q_linear.weight": rnd(E, E, scale=E ** -0.5), "bbox_attention.q_linear.bias": rnd(E, scale=0.1),
          "bbox_attention.k_linear.weight": rnd(E, E, scale=E ** -0.5), "bbox_attention.k_linear.bias": rnd(E, scale=0.1)}
    cins = [E + NH, E + NH, 128, 64, 32]
    couts = [E + NH, 128, 64, 32, 16]
    for i, (ci, co) in enumerate(zip(cins, couts), start=1):
        sd[f"mask_head.lay{i}.weight"] = rnd(co, ci, 3, 3, scale=(9 * ci) ** -0.5 * 1.4)
        sd[f"mask_head.lay{i}.bias"] = rnd(co, scale=0.1)
        sd[f"mask_head.gn{i}.weight"] = rnd(co, scale=0.2) + 1
        sd[f"mask_head.gn{i}.bias"] = rnd(co, scale=0.1)
    sd["mask_head.out_lay.weight"] = rnd(1, 16, 3, 3, scale=(9 * 16) ** -0.5)
    sd["mask_head.out_lay.bias"] = rnd(1, scale=0.1)
    for i, (ci, co) in enumerate(zip((1024, 512, 256), (128, 64, 32)), start=1):
        sd[f"mask_head.adapter{i}.weight"] = rnd(co, ci, 1, 1, scale=ci ** -0.5)
        sd[f"mask_head.adapter{i}.bias"] = rnd(co, scale=0.1)
    # shadows as ShadowBank lays them out: Linear [N, K]; conv OHWI; 1x1 adapters [Cout, Cin]; vectors fp32
    shadow = {}
    for k, v in sd.items():
        if v.dim() == 4 and v.shape[-1] == 3:
            shadow[k] = v.permute(0, 2, 3, 1).contiguous().to(BF)
        elif v.dim() == 4:
            shadow[k] = v.flatten(1).contiguous().to(BF)
        elif v.dim() == 2:
            shadow[k] = v.contiguous().to(BF)
        else:
            shadow[k] = v.contiguous()
    hs = rnd(L, Q * B, E)
    mem = rnd(S, B, E)
    src = rnd(hw, B, E)
    c4, c3, c2 = (torch.relu(rnd(B, h * m, w * m, c)) for m, c in ((2, 1024), (4, 512), (8, 256)))
    small = torch.zeros(B, h, w, dtype=torch.uint8, device=dev)
    small[1, :, w - 2:] = 1
    stage = SimpleNamespace(nheads=NH)
    call = SimpleNamespace(stage=stage, w=shadow, save=True, req=set(sd), n_dec_layers=L, seq_len=S)
    (pred,), saved = MH.mask_fwd(call, hs.to(BF), mem, src.to(BF), c4.to(BF), c3.to(BF), c2.to(BF), small)
    dpred = rnd(*pred.shape)
    (d_hs, d_mem, d_src, d_c4, d_c3, d_c2, _), grads = MH.mask_bwd(call, saved, (True,) * 6 + (False,), dpred)

    import torch.nn.functional as F

    def st(t):  # the kernels store this tensor in bf16: round the value, pass the gradient straight through
        return t + (t.to(BF).float() - t).detach()

    def run_ref(store):
        """store=identity: the oracle's fp32 restatement.  store=st: the same arithmetic with every tensor the kernels
        keep in bf16 rounded at the same place, so that no ReLU mask can flip (tests/test_gpu_blocks.py header)."""
        ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        ins = [hs.clone().requires_grad_(True), mem.clone().requires_grad_(True), src.clone().requires_grad_(True)]
        f_r = [t.permute(0, 3, 1, 2).contiguous().requires_grad_(True) for t in (c4, c3, c2)]
        q_in = ins[0][-1].view(Q, B, E).transpose(0, 1)                          # [B, Q, E]
        memory = ins[1][:hw].permute(1, 2, 0).reshape(B, E, h, w)
        src_nchw = ins[2].permute(1, 2, 0).reshape(B, E, h, w)
        if store is None:
            bbox = O.attention_map(q_in, memory, small.bool(), ref, "bbox_attention.", NH)
            seg = O.mask_head(src_nchw, bbox, f_r, ref, "mask_head.")
        else:
            p = "bbox_attention."
            qq = store(F.linear(q_in, ref[p + "q_linear.weight"], ref[p + "q_linear.bias"]))
            kk = store(F.conv2d(store(memory), ref[p + "k_linear.weight"][:, :, None, None], ref[p + "k_linear.bias"]))
            dh = E // NH
            wts = torch.einsum("bqnc,bnchw->bqnhw", qq.view(B, Q, NH, dh) * (float(dh) ** -0.5), kk.view(B, NH, dh, h, w))
            wts = wts.masked_fill(small.bool()[:, None, None], float("-inf"))
            bbox = store(F.softmax(wts.flatten(3), dim=-1).view_as(wts))
            p = "mask_head."

            def expand(t, n):
                return t.unsqueeze(1).repeat(1, int(n), 1, 1, 1).flatten(0, 1)

            def block(x, i):
                z = store(F.conv2d(x, ref[f"{p}lay{i}.weight"], ref[f"{p}lay{i}.bias"], padding=1))
                return store(F.relu(F.group_norm(z, 8, ref[f"{p}gn{i}.weight"], ref[f"{p}gn{i}.bias"])))

            x = torch.cat([expand(src_nchw, Q), bbox.flatten(0, 1)], 1)
            x = block(block(x, 1), 2)
            for i, fpn in enumerate(f_r, start=1):
                cur = store(F.conv2d(fpn, ref[f"{p}adapter{i}.weight"], ref[f"{p}adapter{i}.bias"]))
                cur = expand(cur, x.shape[0] // cur.shape[0])
                x = store(cur + F.interpolate(x, size=cur.shape[-2:], mode="nearest"))
                x = block(x, i + 2)
            seg = F.conv2d(x, ref[p + "out_lay.weight"], ref[p + "out_lay.bias"], padding=1)
        pr = seg.view(B, Q, seg.shape[-2], seg.shape[-1])
        pr.backward(dpred)
        return pr, ref, ins, f_r

    # Gradients that pass the softmax backward (hs, encoder memory, q_linear / k_linear) see dP = channels E.. of dx0,
    # which is stored in bf16, inside the cancelling term P * (dP - sum P dP): their budget is wider (atol).
    # Parameter gradients of the five conv + GroupNorm + ReLU stages: the bf16-stored conv outputs of the two sides can
    # differ by one bf16 ulp, which flips the ReLU decision of the ~0.1 % of elements whose normalised value is within
    # that ulp of zero.  A flipped element adds or removes a whole dy term from sign-cancelling sums (bias / beta
    # gradients are sums of ~27 k sign-mixed terms: sqrt(2 * 0.1 %) = 4-6 %), while gradients weighted by xhat (gn
    # gamma) and the last layer stay at 0.5 % - measured 4.7e-2 .. 6.5e-2 with `store`, 8e-2 .. 1.0e-1 without.
    for store, ftol, tol, atol in ((None, 1e-2, 0.15, 0.15), (st, 5e-3, 8e-2, 8e-2)):
        pr, ref, (hs_r, mem_r, src_r), f_r = run_ref(store)
        assert pred.shape == pr.shape
        assert rel_err(pred, pr) < ftol
        assert rel_err(d_hs.float(), hs_r.grad) < atol
        assert rel_err(d_mem, mem_r.grad) < atol
        assert rel_err(d_src.float(), src_r.grad) < atol  # sum over the Q queries of bf16-stored, sign-mixed dx0 rows
        for got, want, nm in ((d_c4, f_r[0], "c4"), (d_c3, f_r[1], "c3"), (d_c2, f_r[2], "c2")):
            assert rel_err(got.float().permute(0, 3, 1, 2), want.grad) < atol, nm  # sums over the Q queries (see d_src)
        bad = []
        for k, p_ in ref.items():
            if k == "bbox_attention.k_linear.bias":
                continue
            e = rel_err(grads[k].reshape(p_.shape), p_.grad)
            if not e < (atol if k.startswith("bbox_attention.") else tol):
                bad.append((k, e))
        assert not bad, (store is not None, bad)
