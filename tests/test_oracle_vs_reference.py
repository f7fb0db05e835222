"""Pins oracle/model.py (our CPU restatement) to the *unmodified* reference modules imported from /root/reference.

Runs in the build container only (the GPU box has no /root/reference): every test carries the `reference` marker and
is skipped there.  The same reference outputs are frozen into tests/golden/ by tools/make_golden.py so that the pin
travels.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import max_err, rel_err
from oracle import model as O
from oracle import shims
from synth import make_batch
from toist_b200.tokenizer import CharTokenizer

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not shims.reference_available(), reason="/root/reference not present")]


@pytest.fixture(scope="module")
def ref():
    torch.set_num_threads(8)
    tok = CharTokenizer()
    models = shims.load_reference(tok)
    args = shims.reference_args(["--backbone", "resnet50"])
    torch.manual_seed(0)
    model, criterion, _, weight_dict = models.build_model(args)
    model.eval()
    from util.misc import NestedTensor  # the reference's own container

    images, mask, captions, targets, pm = make_batch(2, 224, 8, seed=7, pad=True)
    with torch.no_grad():
        mc = model(NestedTensor(images, mask), captions, encode_and_save=True)
        out = model(NestedTensor(images, mask), captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, targets, pm, None)
    return dict(model=model, criterion=criterion, mc=mc, out=out, losses=losses, tok=tok,
                batch=(images, mask, captions, targets, pm), args=args)


def _oracle_run(ref):
    images, mask, captions, targets, pm = ref["batch"]
    sd = {k: v.detach() for k, v in ref["model"].state_dict().items()}
    cfg = O.Config(backbone="resnet50")
    tokd = ref["tok"](captions)
    with torch.no_grad():
        mc = O.encode(sd, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
        out = O.decode(sd, cfg, mc)
        losses, idx = O.criterion(cfg, out, tokd, targets, pm)
    return mc, out, losses, idx


def test_forward_matches_reference(ref):
    mc, out, losses, idx = _oracle_run(ref)
    rmc, rout = ref["mc"], ref["out"]
    for k in ("text_memory_resized", "img_memory", "pos_embed", "query_embed"):
        assert rel_err(mc[k], rmc[k]) < 2e-5, k
    assert torch.equal(mc["mask"], rmc["mask"])
    for k in ("pred_logits", "pred_boxes", "proj_queries", "proj_tokens"):
        assert rel_err(out[k], rout[k]) < 5e-5, k
    for l, aux in enumerate(rout["aux_outputs"]):
        for k in ("pred_logits", "pred_boxes", "proj_queries"):
            assert rel_err(out["aux_outputs"][l][k], aux[k]) < 5e-5, (l, k)


def test_losses_and_indices_match_reference(ref):
    _, out, losses, idx = _oracle_run(ref)
    rl = ref["losses"]
    assert set(losses) == set(rl)
    for k, v in rl.items():
        assert abs(float(losses[k]) - float(v)) <= 2e-4 * max(1.0, abs(float(v))), (k, float(losses[k]), float(v))
    # indices: run the reference matcher on the reference outputs, layer by layer
    images, mask, captions, targets, pm = ref["batch"]
    rout = ref["out"]
    layers = list(rout["aux_outputs"]) + [rout]
    oidx = idx[1:] + idx[:1]  # oracle returns [main, aux_0, ...]
    for l, o in enumerate(layers):
        ridx = ref["criterion"].matcher(o, targets, pm)
        for (r0, c0), (r1, c1) in zip(ridx, oidx[l]):
            assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist(), l


def test_segmentation_branch_matches_reference():
    """DETRsegm (models/segmentation.py:40-168) + loss_masks (models/mdetr.py:827-853): the recipe of BASELINE config 3
    (`--mask_model smallconv --frozen_weights --no_aux_loss --no_contrastive_align_loss`) on a small ragged batch."""
    tok = CharTokenizer()
    models = shims.load_reference(tok)
    args = shims.reference_args(["--backbone", "resnet50", "--mask_model", "smallconv", "--frozen_weights", "unused",
                                 "--no_aux_loss", "--no_contrastive_align_loss"])
    torch.manual_seed(0)
    model, criterion, _, weight_dict = models.build_model(args)
    model.eval()
    from util.misc import NestedTensor

    images, mask, captions, targets, pm = make_batch(2, 128, 8, seed=7, pad=True, masks=True)
    with torch.no_grad():
        rmc = model(NestedTensor(images, mask), captions, encode_and_save=True)
        rout = model(NestedTensor(images, mask), captions, encode_and_save=False, memory_cache=rmc)
        rl = criterion(rmc, rout, targets, pm, None)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    cfg = O.Config(backbone="resnet50", prefix="detr.", aux_loss=False, contrastive_align_loss=False)
    tokd = tok(captions)
    with torch.no_grad():
        mc = O.encode(sd, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
        out = O.decode(sd, cfg, mc)
        out["pred_masks"] = O.decode_masks(sd, cfg, mc, out)
        losses, _ = O.criterion(cfg, out, tokd, targets, pm, masks=True)
    assert out["pred_masks"].shape == rout["pred_masks"].shape == (2, 100, 32, 32)
    assert rel_err(out["pred_masks"], rout["pred_masks"]) < 1e-5
    assert set(losses) == set(rl)
    for k, v in rl.items():
        assert abs(float(losses[k]) - float(v)) <= 2e-4 * max(1.0, abs(float(v))), (k, float(losses[k]), float(v))


def test_kmeans_restatement_matches_reference():
    """models/kmeans.py:21-133 (vendored in the reference): same centres, assignments and predictions."""
    shims.load_reference(CharTokenizer())
    from models.kmeans import kmeans as ref_kmeans, kmeans_predict as ref_predict  # reference

    torch.manual_seed(2)
    X = torch.randn(200, 16)
    X[:70] += 2
    for full in (0, 1):
        init = X[:4].clone() + 0.1
        np.random.seed(9)
        rc, rcen = ref_kmeans(X=X, init_cluster_centers=init.clone(), num_clusters=4, full_label=full)
        oc, ocen = O.kmeans(X, init.clone(), 4, full, rng=np.random.RandomState(9))
        assert torch.equal(rc, oc) and max_err(ocen, rcen) < 1e-6
        q = torch.randn(9, 16)
        assert torch.equal(ref_predict(q, rcen), O.kmeans_predict(q, ocen))


def test_matcher_cost_matches_reference_ops(ref):
    from util import box_ops  # reference

    g = torch.Generator().manual_seed(3)
    a = torch.rand(50, 4, generator=g) * 0.4 + 0.1
    b = torch.rand(7, 4, generator=g) * 0.4 + 0.1
    r = box_ops.generalized_box_iou(box_ops.box_cxcywh_to_xyxy(a), box_ops.box_cxcywh_to_xyxy(b))
    o = O.pairwise_giou(O.box_cxcywh_to_xyxy(a), O.box_cxcywh_to_xyxy(b))
    assert torch.equal(r, o)


def test_lsap_restatement_matches_scipy():
    from scipy.optimize import linear_sum_assignment

    rng = np.random.RandomState(1)
    for shape in [(100, 3), (3, 100), (20, 20), (1, 5), (6, 0), (0, 0), (40, 64)]:
        for trial in range(4):
            c = rng.rand(*shape)
            if trial == 1 and c.size:
                c = np.round(c * 3) / 3
            if trial == 2 and c.size:
                c = np.zeros(shape)
            r0, c0 = linear_sum_assignment(c)
            r1, c1 = O.lsap(c)
            assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist(), (shape, trial)
    with pytest.raises(ValueError):
        O.lsap(np.array([[np.nan, 0.0]]))
    with pytest.raises(ValueError):
        O.lsap(np.full((2, 2), np.inf))


def test_lr_schedule_matches_reference():
    """toist_b200.util.optim.adjust_learning_rate against the reference's util/optim.py:29-90 for all four schedules."""
    import argparse
    import importlib.util
    import random

    from toist_b200.util import optim as ours

    spec = importlib.util.spec_from_file_location("_ref_optim", "/root/reference/util/optim.py")
    refmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refmod)

    class Opt:
        def __init__(self):
            self.param_groups = [{"lr": 0.0}, {"lr": 0.0}, {"lr": 0.0}]

    rng = random.Random(0)
    for schedule in ("step", "multistep", "linear_with_warmup", "all_linear_with_warmup"):
        for _ in range(50):
            a = argparse.Namespace(fraction_warmup_steps=rng.choice([0.0, 0.01, 0.1]), schedule=schedule,
                                   lr_drop=rng.choice([5, 10, 35]), epochs=rng.choice([30, 120, 200]), lr=1e-4,
                                   lr_backbone=1e-5, text_encoder_lr=5e-5)
            total = rng.choice([1, 100, 12345])
            epoch, step = rng.randrange(0, a.epochs), rng.randrange(0, total + 1)
            o1, o2 = Opt(), Opt()
            refmod.adjust_learning_rate(o1, epoch, step, total, a)
            ours.adjust_learning_rate(o2, epoch, step, total, a)
            assert o1.param_groups == o2.param_groups, (schedule, epoch, step, total)
            # distillation recipe: student + noun model, six groups (util/optim.py:92-152)
            d1, d2 = Opt(), Opt()
            d1.param_groups = [{"lr": 0.0} for _ in range(6)]
            d2.param_groups = [{"lr": 0.0} for _ in range(6)]
            refmod.dis_adjust_learning_rate(d1, epoch, step, total, a)
            ours.dis_adjust_learning_rate(d2, epoch, step, total, a)
            assert d1.param_groups == d2.param_groups, (schedule, epoch, step, total)


def test_gradients_match_reference(ref):
    """The oracle's backward is pinned too: gradient of the weighted loss sum (engine.py:72,88) w.r.t. EVERY trainable
    tensor, reference vs oracle.  In particular loss_contrastive_align is differentiated (models/mdetr.py:601-666 has
    no @torch.no_grad; weight 1 per layer at :1068-1069): both contrastive projections receive a gradient."""
    images, mask, captions, targets, pm = ref["batch"]
    model, criterion = ref["model"], ref["criterion"]
    from util.misc import NestedTensor  # reference

    weight_dict = shims.load_reference(ref["tok"]).build_model(ref["args"])[3]
    model.zero_grad(set_to_none=True)
    mc = model(NestedTensor(images, mask), captions, encode_and_save=True)
    out = model(NestedTensor(images, mask), captions, encode_and_save=False, memory_cache=mc)
    losses = criterion(mc, out, targets, pm, None)
    assert losses["loss_contrastive_align"].requires_grad
    total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)
    total.backward()
    rgrads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    assert float(rgrads["contrastive_align_projection_image.weight"].norm()) > 0
    assert float(rgrads["contrastive_align_projection_text.weight"].norm()) > 0

    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k in trainable:
        sd[k].requires_grad_(True)
    cfg = O.Config(backbone="resnet50")
    tokd = ref["tok"](captions)
    omc = O.encode(sd, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
    oout = O.decode(sd, cfg, omc)
    olosses, _ = O.criterion(cfg, oout, tokd, targets, pm)
    ototal = sum(olosses[k] * weight_dict[k] for k in olosses if k in weight_dict)
    ototal.backward()
    assert abs(float(ototal.detach()) - float(total.detach())) <= 2e-4 * abs(float(total.detach()))
    missing = [k for k in rgrads if sd[k].grad is None]
    assert not missing, missing
    extra = [k for k in trainable if sd[k].grad is not None and k not in rgrads]
    assert not extra, extra
    # RoBERTa key biases: the softmax is invariant to them, their exact gradient is zero and what is left is rounding
    errs = sorted(((rel_err(sd[k].grad, rgrads[k]), k) for k in rgrads
                   if float(rgrads[k].norm()) > 0 and "key.bias" not in k), reverse=True)
    assert errs[0][0] < 2e-3, errs[:5]


def test_ragged_captions_gradients_match_reference(ref):
    """Captions of different lengths: the shorter one is padded with <pad> (id 1).  RoBERTa's word and position tables
    are nn.Embedding(padding_idx=1): their <pad> rows receive NO gradient, and padded keys are masked in the text
    encoder and in the cross-modal encoder.  Reference vs oracle, embedding gradients and a few others."""
    from synth import make_batch as mk
    from util.misc import NestedTensor  # reference

    model, criterion = ref["model"], ref["criterion"]
    weight_dict = shims.load_reference(ref["tok"]).build_model(ref["args"])[3]
    images, mask, captions, targets, pm = mk(2, 128, 12, seed=9, pad=True)
    captions = [captions[0], captions[1][4:]]  # 10 and 6 characters -> 12 and 8 tokens
    targets[1]["tokens_positive"] = [[[0, len(captions[1])]] for _ in targets[1]["tokens_positive"]]
    tokd = ref["tok"](captions)
    assert int((tokd["input_ids"] == 1).sum()) == 4
    model.zero_grad(set_to_none=True)
    mc = model(NestedTensor(images, mask), captions, encode_and_save=True)
    out = model(NestedTensor(images, mask), captions, encode_and_save=False, memory_cache=mc)
    losses = criterion(mc, out, targets, pm, None)
    sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict).backward()
    names = ["transformer.text_encoder.embeddings.word_embeddings.weight",
             "transformer.text_encoder.embeddings.position_embeddings.weight",
             "transformer.text_encoder.encoder.layer.0.attention.self.value.weight",
             "transformer.encoder.layers.0.linear1.weight", "contrastive_align_projection_text.weight", "input_proj.weight"]
    lookup = dict(model.named_parameters())
    rg = {n: lookup[n].grad.clone() for n in names}
    model.zero_grad(set_to_none=True)
    assert float(rg[names[0]][1].abs().max()) == 0.0 and float(rg[names[1]][1].abs().max()) == 0.0  # <pad> rows
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k in names:
        sd[k].requires_grad_(True)
    cfg = O.Config(backbone="resnet50")
    omc = O.encode(sd, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
    oout = O.decode(sd, cfg, omc)
    olosses, _ = O.criterion(cfg, oout, tokd, targets, pm)
    sum(olosses[k] * weight_dict[k] for k in olosses if k in weight_dict).backward()
    for k, v in losses.items():
        assert abs(float(olosses[k].detach()) - float(v.detach())) <= 2e-4 * max(1.0, abs(float(v.detach()))), k
    for n in names:
        assert rel_err(sd[n].grad, rg[n]) < 2e-3, (n, rel_err(sd[n].grad, rg[n]))
