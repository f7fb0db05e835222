"""GPU parity against the golden vectors frozen from the unmodified reference (tools/make_golden.py).

* matcher: identical fp32 inputs -> cost within 2e-6 abs, assignment indices BIT-IDENTICAL (north_star).
* criterion: the reference's own fp32 outputs of BASELINE config 1 fed to our CUDA SetCriterion -> all 30 loss terms
  within 1e-4 relative and the assignments of all 6 decoder layers bit-identical.
* model: our CUDA forward from the same seed-0 weights -> outputs / losses within the bf16 budget stated below.
"""
from __future__ import annotations

from pathlib import Path

import pytest
import torch

from conftest import max_err, rel_err, strided_sample
from toist_b200.synth import make_args, make_batch, targets_to

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
DEV = "cuda"


@pytest.fixture(scope="module")
def config1_gold():
    return torch.load(GOLD / "config1_r50.pt", weights_only=False)


def test_matcher_bit_identical_to_reference_goldens():
    from toist_b200.models.matcher import HungarianMatcher, match_layers, pack_targets

    g = torch.load(GOLD / "matcher_cases.pt", weights_only=False)
    m = HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)
    for c in g["cases"]:
        targets = [{"boxes": b.to(DEV), "labels": torch.ones(len(b), dtype=torch.long, device=DEV)} for b in c["tgt_boxes"]]
        out = {"pred_logits": c["logits"].to(DEV), "pred_boxes": c["boxes"].to(DEV)}
        idx = m(out, targets, c["positive_map"].to(DEV))
        for (r0, c0), (r1, c1) in zip(idx, c["indices"]):
            assert r0.dtype == torch.int64 and r0.device.type == "cpu"
            assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist()
        pt = pack_targets(targets, c["positive_map"].to(DEV), DEV)
        _, _, cost = match_layers(out["pred_logits"][None].contiguous(), out["pred_boxes"][None].contiguous(), pt, 1.0,
                                  5.0, 2.0)
        off = 0
        for b, n in enumerate(pt.counts):
            if n:
                assert max_err(cost[0, b, :, :n], c["cost"][b][:, off:off + n]) < 2e-6
            off += n


def test_criterion_on_reference_outputs(config1_gold):
    """Identical fp32 inputs (the reference's own predictions): losses to 1e-4, assignments bit-identical."""
    from toist_b200.models import build_model
    from toist_b200.tokenizer import CharTokenizer

    g = config1_gold
    _, criterion, _, _ = build_model(make_args("resnet50"))
    b = g["batch"]
    _, _, captions, targets, pm = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"])
    tok = CharTokenizer()(captions)
    L = g["pred_logits"].shape[0]
    layers = [{"pred_logits": g["pred_logits"][l].to(DEV), "pred_boxes": g["pred_boxes"][l].to(DEV),
               "proj_queries": g["proj_queries"][l].to(DEV), "proj_tokens": g["proj_tokens"].to(DEV), "tokenized": tok}
              for l in range(L)]
    outputs = dict(layers[-1])
    outputs["aux_outputs"] = layers[:-1]
    losses = criterion({}, outputs, targets_to(targets, DEV), pm.to(DEV), None)
    assert list(losses) == list(g["losses"]), "loss dictionary keys / order differ from the reference"
    for k, v in g["losses"].items():
        assert abs(float(losses[k]) - v) <= 1e-4 * max(1.0, abs(v)), (k, float(losses[k]), v)
    idx = criterion.last_indices()
    for l in range(L):
        for (r0, c0), (r1, c1) in zip(idx[l], g["indices"][l]):
            assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist(), l


def test_model_forward_against_reference_golden(config1_gold):
    """Same seed-0 weights as the reference (checksums verified), bf16 tensor-core forward vs the reference's fp32
    forward.  Budget: 3e-2 norm-wise on every output tensor and 2e-2 relative on every loss term; torch's own bf16
    autocast of the same network sits at 1e-2 .. 1.6e-2 on these tensors (tests/e2e_report.py --calibrate), i.e. the
    distance is the number format, not the kernels.  north_star's 1e-3 holds per kernel on identical inputs
    (tools/gemm_selftest.py, test_gpu_ops.py) and for every fp32 stage (matcher, criterion: tests above)."""
    from toist_b200.models import build_model
    from toist_b200.util.misc import NestedTensor

    g = config1_gold
    torch.manual_seed(0)
    model, criterion, _, _ = build_model(make_args("resnet50"))
    for k, v in g["state_checksum"].items():
        assert abs(float(model.state_dict()[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), k
    model.cuda().eval()
    b = g["batch"]
    images, mask, captions, targets, pm = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"])
    samples = NestedTensor(images.to(DEV), mask.to(DEV))
    with torch.no_grad():
        mc = model(samples, captions, encode_and_save=True)
        out = model(samples, captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, targets_to(targets, DEV), pm.to(DEV), None)
    assert torch.equal(mc["mask"].cpu(), g["mask"])
    assert rel_err(mc["pos_embed"][:, 0], g["pos_embed_row0"]) < 1e-6
    assert rel_err(mc["img_memory"], g["img_memory"].float()) < 3e-2
    assert rel_err(mc["text_memory_resized"], g["text_memory_resized"].float()) < 3e-2
    st = out["_b200_stacked"]
    for k in ("pred_logits", "pred_boxes", "proj_queries", "proj_tokens"):
        assert rel_err(st[k], g[k]) < 3e-2, k
    assert out["pred_logits"].shape == (2, 100, 256) and out["pred_boxes"].shape == (2, 100, 4)
    assert len(out["aux_outputs"]) == 5
    assert list(losses) == list(g["losses"])
    for k, v in g["losses"].items():
        assert abs(float(losses[k]) - v) <= 2e-2 * max(1.0, abs(v)), (k, float(losses[k]), v)
    # calibrated: the unmodified reference under torch's own bf16 autocast is this far from its fp32 self on the same
    # batch (tools/make_golden.py: img_memory 2.4e-2, pred_logits 1.5e-2, proj_tokens 1.6e-2); we must not be further
    fl = g["bf16_forward_floor"]
    assert rel_err(mc["img_memory"], g["img_memory"].float()) < 1.25 * fl["img_memory"]
    for k in ("pred_logits", "proj_queries", "proj_tokens"):
        last = st[k][-1] if k != "proj_tokens" else st[k]
        ref = g[k][-1] if k != "proj_tokens" else g[k]
        assert rel_err(last, ref) < 1.25 * fl[k], (k, rel_err(last, ref), fl[k])


def test_criterion_gradients_on_reference_outputs(config1_gold):
    """fp32 in, fp32 out: d(weighted loss sum)/d(pred_logits, pred_boxes, proj_queries, proj_tokens) of our CUDA
    SetCriterion on the reference's own predictions vs the oracle (whose backward is pinned to the reference by
    tests/test_oracle_vs_reference.py::test_gradients_match_reference).  Covers loss_contrastive_align's gradient
    (models/mdetr.py:601-666), which the reference back-propagates with weight 1 per layer (:1068-1069)."""
    from oracle import model as O
    from toist_b200.models import build_model
    from toist_b200.tokenizer import CharTokenizer

    g = config1_gold
    _, criterion, _, wd = build_model(make_args("resnet50"))
    b = g["batch"]
    _, _, captions, targets, pm = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"])
    tok = CharTokenizer()(captions)
    L = g["pred_logits"].shape[0]
    names = ("pred_logits", "pred_boxes", "proj_queries")

    def leaves(dev):
        st = {k: g[k].to(dev).clone().requires_grad_(True) for k in names}
        ptok = g["proj_tokens"].to(dev).clone().requires_grad_(True)
        layers = [{**{k: st[k][l] for k in names}, "proj_tokens": ptok, "tokenized": tok} for l in range(L)]
        outputs = dict(layers[-1])
        outputs["aux_outputs"] = layers[:-1]
        return st, ptok, outputs

    st, ptok, outputs = leaves(DEV)
    losses = criterion({}, outputs, targets_to(targets, DEV), pm.to(DEV), None)
    assert losses["loss_contrastive_align"].requires_grad and not losses["cardinality_error"].requires_grad
    sum(losses[k] * wd[k] for k in losses if k in wd).backward()
    ost, optok, ooutputs = leaves("cpu")
    olosses, _ = O.criterion(O.Config(backbone="resnet50"), ooutputs, tok, targets, pm)
    sum(olosses[k] * wd[k] for k in olosses if k in wd).backward()
    for k in names:
        assert rel_err(st[k].grad, ost[k].grad) < 1e-4, (k, rel_err(st[k].grad, ost[k].grad))
    assert float(optok.grad.norm()) > 0
    assert rel_err(ptok.grad, optok.grad) < 1e-4, rel_err(ptok.grad, optok.grad)


def test_model_gradients_against_reference_golden(config1_gold):
    """Detection backward pinned to the REFERENCE: same seed-0 weights, same batch, the reference's own assignments
    forced into our criterion (bf16 noise flips near-tie costs at random init), gradient of the weighted loss sum w.r.t.
    all 440 trainable tensors vs the gradients the unmodified reference produced (tools/make_golden.py: L2 norm and a
    2048-element strided sample of every tensor, full tensors for the head-side parameters).  Budget: the bf16
    activation format (see test_gpu_model.py::test_gradients_at_the_bf16_noise_floor for the calibration)."""
    from toist_b200.models import build_model
    from toist_b200.util.misc import NestedTensor

    g = config1_gold
    torch.manual_seed(0)
    model, criterion, _, wd = build_model(make_args("resnet50"))
    model.cuda().eval()
    b = g["batch"]
    images, mask, captions, targets, pm = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"])
    samples = NestedTensor(images.to(DEV), mask.to(DEV))
    criterion.force_match(g["indices"])
    mc = model(samples, captions, encode_and_save=True)
    out = model(samples, captions, encode_and_save=False, memory_cache=mc)
    losses = criterion(mc, out, targets_to(targets, DEV), pm.to(DEV), None)
    total = sum(losses[k] * wd[k] for k in losses if k in wd)
    total.backward()
    criterion.force_match(None)
    assert abs(float(total.detach()) - g["total"]) <= 2e-2 * abs(g["total"])
    gold = g["grads"]
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert set(grads) == set(gold["norm"]), sorted(set(grads) ^ set(gold["norm"]))[:10]
    # head-side parameters, full tensors: the contrastive projections train (reference: norm 1.9 / 0.9)
    full = {k: rel_err(grads[k], v) for k, v in gold["full"].items()}
    table = []
    for k, ref_norm in gold["norm"].items():
        if "key.bias" in k or ref_norm < 1e-12:  # exact gradient is zero (softmax is invariant to a key bias)
            continue
        ours = grads[k]
        table.append((rel_err(strided_sample(ours), gold["sample"][k]), float(ours.double().norm()) / ref_norm, k))
    out_dir = Path(__file__).resolve().parent.parent / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    with open(out_dir / "grad_parity_config1.txt", "w") as f:
        for e, r, k in sorted(table, reverse=True):
            f.write(f"{e:.4e} norm_ratio {r:.4f} {k}\n")
        for k, e in full.items():
            f.write(f"full {e:.4e} {k}\n")
    # Budget = the number format, measured on the REFERENCE ITSELF: gold["bf16_floor"][k] is the distance of the
    # unmodified reference's gradient under torch's bf16 autocast from its own fp32 gradient (same assignments, same
    # strided sample).  Ill-conditioned tensors (the contrastive projections differentiate a softmax over nearly
    # identical random-init token embeddings at temperature 0.07: 0.36 for torch, 0.34 for us) stay ill-conditioned
    # for any bf16 run; ours must stay within 2.5x of torch's per tensor (measured max 1.7x) and be no worse overall
    # (measured median ratio 0.85).
    floor = gold["bf16_floor"]
    ratios = sorted((e / max(floor[k], 1e-12), e, floor[k], k) for e, _, k in table if k in floor)
    assert len(ratios) > 400
    bad = [r for r in ratios if r[1] > max(2.5 * r[2], 2e-2)]
    assert not bad, bad[-5:]
    assert ratios[len(ratios) // 2][0] < 1.2, ratios[len(ratios) // 2]
    for k in ("contrastive_align_projection_image.weight", "contrastive_align_projection_text.weight"):
        assert float(grads[k].norm()) > 0 and full[k] < 2.5 * floor[k], (k, full[k], floor[k])
    for k in ("class_embed.weight", "bbox_embed.layers.2.weight", "transformer.decoder.norm.weight"):
        assert full[k] < 3e-2, (k, full[k])
    nr = [r for _, r, _ in table]
    assert 0.7 < min(nr) and max(nr) < 1.5, (min(nr), max(nr))


def test_matcher_1000_problems_bit_identical_to_scipy_on_our_cost():
    """>= 1000 random seeds + degenerate cases.  125 batches x 8 images: the device matcher (cost kernel +
    warp-parallel LSAP) against the oracle (reference operation order, models/matcher.py:63-87, + scipy's
    linear_sum_assignment when it is installed, else the pinned restatement) on identical fp32 inputs.  Includes images
    without targets, duplicated queries (exact ties), nearly identical queries (near ties at the 1e-7 level) and
    T close to Q.  Assignments must be identical; the cost matrix within 2e-6."""
    from oracle import model as O
    from toist_b200.models.matcher import HungarianMatcher

    m = HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)
    g = torch.Generator().manual_seed(2024)
    n_problems = n_ties = 0
    for it in range(125):
        B, Q = 8, (100 if it % 5 else 16)
        logits = torch.randn(B, Q, 256, generator=g) * (0.1 if it % 3 == 0 else 2.0)  # 0.1: random-init-like near ties
        boxes = torch.cat([torch.rand(B, Q, 2, generator=g) * 0.5 + 0.25, torch.rand(B, Q, 2, generator=g) * 0.3 + 0.05], -1)
        if it % 7 == 0:   # exact ties: half of the queries are copies of query 0
            logits[:, Q // 2:] = logits[:, :1]
            boxes[:, Q // 2:] = boxes[:, :1]
            n_ties += 1
        if it % 11 == 0:  # near ties: copies perturbed in the last bits
            logits[:, 1::2] = logits[:, 0::2] * (1 + 1e-7)
            boxes[:, 1::2] = boxes[:, 0::2]
        counts = [int(torch.randint(0, 6, (1,), generator=g)) for _ in range(B)]
        if it % 13 == 0:
            counts[0] = min(Q - 1, 14)  # many targets
        targets = []
        for n in counts:
            tb = torch.cat([torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.3 + 0.05], -1)
            targets.append({"boxes": tb, "labels": torch.ones(n, dtype=torch.long)})
        T = sum(counts)
        pm = torch.zeros(T, 256)
        for r in range(T):
            lo = 1 + int(torch.randint(0, 12, (1,), generator=g))
            pm[r, lo: lo + 1 + int(torch.randint(0, 4, (1,), generator=g))] = 1
        pm = pm / (pm.sum(-1, keepdim=True) + 1e-6)
        want = O.hungarian_match(logits, boxes, targets, pm, 1.0, 5.0, 2.0)
        got = m({"pred_logits": logits.to(DEV), "pred_boxes": boxes.to(DEV)},
                [{k: v.to(DEV) for k, v in t.items()} for t in targets], pm.to(DEV))
        for b, ((r0, c0), (r1, c1)) in enumerate(zip(got, want)):
            assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist(), (it, b)
            n_problems += 1
    assert n_problems == 1000 and n_ties >= 15
