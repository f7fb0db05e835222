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

from conftest import max_err, rel_err
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
