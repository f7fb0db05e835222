"""CPU suite: pins the oracle to the golden vectors frozen from the unmodified reference (tools/make_golden.py),
checks the host-side logic, and that libtoist_b200.so loads and exports the whole C ABI (no GPU compute here)."""
from __future__ import annotations

import ctypes
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import max_err, rel_err
from oracle import model as O
from toist_b200.synth import make_args, make_batch

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def matcher_gold():
    return torch.load(GOLD / "matcher_cases.pt", weights_only=False)


@pytest.fixture(scope="module")
def config1_gold():
    return torch.load(GOLD / "config1_r50.pt", weights_only=False)


def test_oracle_matcher_reproduces_reference_goldens(matcher_gold):
    for c in matcher_gold["cases"]:
        targets = [{"boxes": b} for b in c["tgt_boxes"]]
        tgt = torch.cat(c["tgt_boxes"])
        cost = O.matcher_cost(c["logits"], c["boxes"], tgt, c["positive_map"], 1.0, 5.0, 2.0)
        assert max_err(cost, c["cost"]) < 1e-6
        idx = O.hungarian_match(c["logits"], c["boxes"], targets, c["positive_map"], 1.0, 5.0, 2.0)
        for (r0, c0), (r1, c1) in zip(idx, c["indices"]):
            assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist()


def test_oracle_model_reproduces_reference_golden(config1_gold):
    """Our parameter containers initialise bit-identically to the reference under the same seed, and the oracle run on
    those weights reproduces the reference's forward, losses and assignments (BASELINE config 1)."""
    from toist_b200.models import build_model

    g = config1_gold
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    torch.manual_seed(0)
    model, _, _, weight_dict = build_model(make_args("resnet50", device="cpu"))
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    assert set(sd) == set(g["state_checksum"])
    for k, v in g["state_checksum"].items():
        assert abs(float(sd[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), k
    assert dict(weight_dict) == g["weight_dict"]
    b = g["batch"]
    images, mask, captions, targets, pm = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"])
    tokd = model.transformer.tokenizer(captions)
    cfg = O.Config(backbone="resnet50")
    with torch.no_grad():
        mc = O.encode(sd, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
        out = O.decode(sd, cfg, mc)
        losses, idx = O.criterion(cfg, out, tokd, targets, pm)
    assert torch.equal(mc["mask"], g["mask"])
    assert rel_err(mc["img_memory"], g["img_memory"].float()) < 1e-3  # golden stored in fp16
    assert rel_err(mc["pos_embed"][:, 0], g["pos_embed_row0"]) < 1e-6
    layers = list(out["aux_outputs"]) + [out]
    for k in ("pred_logits", "pred_boxes", "proj_queries"):
        assert rel_err(torch.stack([o[k] for o in layers]), g[k]) < 1e-4, k
    assert rel_err(out["proj_tokens"], g["proj_tokens"]) < 1e-4
    assert set(losses) == set(g["losses"])
    for k, v in g["losses"].items():
        assert abs(float(losses[k]) - v) <= 2e-4 * max(1.0, abs(v)), k
    oidx = idx[1:] + idx[:1]
    for l, layer in enumerate(g["indices"]):
        for (r0, c0), (r1, c1) in zip(oidx[l], layer):
            assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist(), l


def test_oracle_mask_branch_reproduces_reference_golden():
    """DETRsegm mask branch + mask losses (BASELINE config 3 recipe, small): same-seed initialisation is bit-identical
    to the reference's and the oracle reproduces pred_masks, the six loss terms and the mask-branch gradients."""
    from toist_b200.models import build_model

    g = torch.load(GOLD / "config3_r50_segm_small.pt", weights_only=False)
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    torch.manual_seed(0)
    args = make_args("resnet50", device="cpu", masks=True, mask_model="smallconv", frozen_weights="unused",
                     aux_loss=False, contrastive_align_loss=False)
    model, _, _, weight_dict = build_model(args)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert set(sd) == set(g["state_checksum"])
    for k, v in g["state_checksum"].items():
        assert abs(float(sd[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), k
    assert dict(weight_dict) == g["weight_dict"]
    trainable = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    assert trainable == sorted(g["grads"])
    b = g["batch"]
    images, mask, captions, targets, pm = make_batch(b["batch"], b["size"], b["tokens"], seed=b["seed"], pad=b["pad"],
                                                     masks=True)
    tokd = model.detr.transformer.tokenizer(captions)
    cfg = O.Config(backbone="resnet50", prefix="detr.", aux_loss=False, contrastive_align_loss=False)
    for k in trainable:
        sd[k].requires_grad_(True)
    mc = O.encode(sd, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
    out = O.decode(sd, cfg, mc)
    out["pred_masks"] = O.decode_masks(sd, cfg, mc, out)
    losses, idx = O.criterion(cfg, out, tokd, targets, pm, masks=True)
    assert rel_err(out["pred_masks"], g["pred_masks"]) < 1e-5
    for k, v in g["losses"].items():
        assert abs(float(losses[k]) - v) <= 2e-4 * max(1.0, abs(v)), k
    for (r0, c0), (r1, c1) in zip(idx[0], g["indices"]):
        assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist()
    sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict).backward()
    for k in trainable:
        if k == "bbox_attention.k_linear.bias":  # softmax is shift invariant: this gradient is rounding noise around 0
            assert float(sd[k].grad.abs().max()) < 1e-12
            continue
        assert rel_err(sd[k].grad, g["grads"][k]) < 1e-4, k


def test_oracle_softkd_reproduces_reference_golden():
    """loss_softkd + softkd_matcher (models/mdetr.py:520-599) on the hand-made golden cases: value and gradient."""
    g = torch.load(GOLD / "config5_softkd_small.pt", weights_only=False)
    for c in g["softkd_cases"]:
        ls = c["logits_sth"].clone().requires_grad_(True)
        Q = ls.shape[1]
        val = O.loss_softkd({"pred_logits": c["logits_noun"], "pred_boxes": c["boxes_noun"]},
                            {"pred_logits": ls, "pred_boxes": c["boxes_sth"]}, c["idx_noun"], c["idx_sth"], Q)
        assert abs(float(val) - c["loss"]) <= 1e-6 * max(1.0, abs(c["loss"]))
        val.backward()
        assert rel_err(ls.grad, c["grad_logits_sth"]) < 1e-5


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    from toist_b200 import _lib

    protos = _lib.parse_header()
    assert len(protos) >= 40
    lib = _lib.load()
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.toist_abi_version() == 1
    assert lib.toist_sizeof_gemm_desc() == ctypes.sizeof(_lib.GemmDesc)
    assert isinstance(lib.toist_last_error(), bytes)


def test_host_lsap_matches_scipy_and_oracle():
    """toist_lsap_f64 is host code (C++), so it runs here: identical to scipy (the reference's solver) and the oracle."""
    from scipy.optimize import linear_sum_assignment

    from toist_b200 import kernels as K

    rng = np.random.RandomState(0)
    for nr, nc in [(100, 4), (4, 100), (97, 97), (1, 1), (5, 0), (0, 5), (30, 1024), (100, 1), (3, 3)]:
        for trial in range(3):
            c = rng.rand(nr, nc)
            if trial == 1 and c.size:
                c = np.round(c * 4) / 4
            if trial == 2 and c.size:
                c = np.zeros((nr, nc))
            r0, c0 = linear_sum_assignment(c)
            r1, c1 = K.lsap_host(c)
            assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist(), (nr, nc, trial)
            if nr * nc <= 400:
                r2, c2 = O.lsap(c)
                assert r0.tolist() == r2.tolist() and c0.tolist() == c2.tolist()
    with pytest.raises(ValueError):
        K.lsap_host(np.array([[np.nan, 1.0], [1.0, 2.0]]))
    with pytest.raises(ValueError):
        K.lsap_host(np.array([[np.inf, np.inf], [np.inf, np.inf]]))
    with pytest.raises(ValueError):
        K.lsap_host(np.array([[-np.inf, 1.0]]))


def test_target_packing_and_token_spans_host_logic():
    from toist_b200.models.matcher import indices_from_match, pack_targets
    from toist_b200.models.mdetr import build_token_positive
    from toist_b200.tokenizer import CharTokenizer

    _, _, captions, targets, pm = make_batch(5, 32, 8, seed=3)
    p = pack_targets(targets, pm, "cpu")
    assert p.counts == (1, 2, 3, 4, 1) and p.t_max == 4
    off = 0
    for b, n in enumerate(p.counts):
        assert torch.equal(p.boxes[b, :n], targets[b]["boxes"])
        assert torch.equal(p.posmap[b, :n], pm[off:off + n])
        assert bool((p.boxes[b, n:] == 0).all()) and bool((p.posmap[b, n:] == 0).all())
        off += n
    assert p.count.tolist() == list(p.counts)
    tok = CharTokenizer()(captions)
    tp = build_token_positive(tok, targets, p.t_max, 8)
    for b, n in enumerate(p.counts):
        assert bool((tp[b, :n, 1:7] == 1).all()) and bool((tp[b, :n, 0] == 0).all()) and bool((tp[b, :n, 7] == 0).all())
        assert bool((tp[b, n:] == 0).all())
    mq = torch.tensor([[5, -1, -1, -1], [9, 2, -1, -1]], dtype=torch.int32)
    idx = indices_from_match(mq, (1, 2))
    assert idx[0][0].tolist() == [5] and idx[0][1].tolist() == [0]
    assert idx[1][0].tolist() == [2, 9] and idx[1][1].tolist() == [1, 0]


def test_no_cpu_fallback_in_product_path():
    """The product path must fail loudly without a CUDA device instead of silently computing on the CPU."""
    from toist_b200.models import build_model
    from toist_b200.util.misc import NestedTensor

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    model, criterion, _, _ = build_model(make_args("resnet50", device="cpu"))
    model.eval()
    images, mask, captions, targets, pm = make_batch(1, 64, 8)
    with pytest.raises(RuntimeError):
        model(NestedTensor(images, mask), captions, encode_and_save=True)
    with pytest.raises(RuntimeError):
        criterion({}, {"pred_logits": torch.zeros(1, 100, 256), "pred_boxes": torch.zeros(1, 100, 4)}, targets, pm, None)


def test_loss_terms_node_matches_indexing():
    """The criterion hands out its [5, L] loss tensor as 5 * L scalars through one autograd node (models/mdetr.py
    _LossTerms): values and the gradient of any weighted sum must equal plain `out[row, l]` indexing, including unused
    and detached entries (cardinality / contrastive rows are logging-only in the reference)."""
    from toist_b200.models.mdetr import _LossTerms

    torch.manual_seed(0)
    out = torch.randn(5, 6, requires_grad=True)
    w = torch.rand(5, 6)
    ref = torch.randn(5, 6).detach().copy_(out.detach()).requires_grad_(True)
    cells = _LossTerms.apply(out)
    assert len(cells) == 30 and all(c.dim() == 0 for c in cells)
    used = [(r, l) for r in range(3) for l in range(6) if (r + l) % 4 != 0]
    total = sum(cells[r * 6 + l] * float(w[r, l]) for r, l in used) + cells[3 * 6 + 2].detach() * 7.0
    total_ref = sum(ref[r, l] * float(w[r, l]) for r, l in used) + ref[3, 2].detach() * 7.0
    assert torch.allclose(total, total_ref, rtol=1e-6, atol=1e-6)
    total.backward()
    total_ref.backward()
    assert torch.allclose(out.grad, ref.grad, rtol=0, atol=1e-7)


def test_h2d_keeps_values_and_is_a_no_op_without_cuda():
    from toist_b200.util.misc import h2d

    t = torch.arange(12, dtype=torch.int64).view(3, 4)
    assert torch.equal(h2d(t, "cpu"), t)
    assert h2d(t, "cpu", torch.float32).dtype == torch.float32


def test_library_contains_tcgen05_and_tma_instructions():
    """The hot kernels are hand-written sm_100a code: the built library must contain tcgen05.mma (SASS UTCHMMA), TMA loads /
    stores (UTMALDG / UTMASTG), TMEM loads (LDTM) and mbarrier operations (SYNCS).  Needs cuobjdump (CUDA toolkit)."""
    import shutil
    import subprocess

    from toist_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists() or not _lib.lib_path().exists():
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run([cuobjdump, "-sass", str(_lib.lib_path())], capture_output=True, text=True, timeout=300).stdout
    for mnemonic, at_least in (("UTCHMMA", 100), ("UTMALDG", 100), ("UTMASTG", 12), ("LDTM", 20), ("SYNCS", 500)):
        assert sass.count(mnemonic) >= at_least, (mnemonic, sass.count(mnemonic))


def test_bench_reference_arm_prints_one_valid_json_line():
    """`bench.py --impl reference` (the oracle port on the host cores, a bounded 2-image sample of the bench workload)
    must print exactly one JSON line with the contract's keys; it needs no GPU."""
    import json
    import subprocess
    import sys

    root = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("images/sec") and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
