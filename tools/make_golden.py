"""Freezes outputs of the UNMODIFIED reference (/root/reference, imported through oracle/shims.py) as small fixtures
under tests/golden/.  Runs in the build container only; the fixtures travel to the GPU box, the reference does not.

    python tools/make_golden.py

Fixtures (torch.save, fp32 unless noted; library versions recorded in each file):
  matcher_cases.pt   HungarianMatcher.forward on hand-made inputs (ragged target counts, an image without targets,
                     exact ties): cost matrix C and the assignment indices            (models/matcher.py:39-87)
  config1_r50.pt     BASELINE config 1: ResNet-50 TOIST, 2 x 3 x 480 x 480, 8-token captions, seed-0 random init,
                     eval mode: memory_cache tensors, all-layer outputs, the 30 loss terms, the assignments of all 6
                     decoder layers, and per-tensor checksums of the state dict          (models/mdetr.py:377-462,990-1021)
  config3_r50_segm_small.pt  the mask recipe of BASELINE config 3 (frozen detector + DETRsegm mask head) on a
                     2 x 3 x 128 x 128 ragged batch: pred_masks, the 6 loss terms, the assignment, and the gradients of
                     every mask-branch parameter                          (models/segmentation.py:40-273, mdetr.py:827-853)
  config5_softkd_small.pt  distillation branch of SetCriterion with soft-KD: fp32 predictions of teacher and student,
                     the 66 loss terms, d(soft-KD)/d(student logits)                        (models/mdetr.py:520-599,887-989)
  nsthl2_cases.pt    SetCriterion.loss_nsthl2 on hand-made teacher / student text memories (incl. an image without targets
                     and the all-empty batch): the loss and d loss / d student text memory     (models/mdetr.py:668-781)
  cluster_cases.pt   ClusterCriterion driven for 10 steps on random features (memory bank, k-means, replacement,
                     cluster-feature loss and its gradient), run on the CPU by patching .cuda()  (models/mdetr.py:29-312)
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import shims  # noqa: E402
from toist_b200.synth import make_batch  # noqa: E402
from toist_b200.tokenizer import CharTokenizer  # noqa: E402

OUT = ROOT / "tests" / "golden"


def versions():
    import scipy
    import torchvision
    import transformers

    return {"torch": torch.__version__, "torchvision": torchvision.__version__, "transformers": transformers.__version__,
            "scipy": scipy.__version__, "reference": "AIR-DISCOVER/TOIST @ e3e17ae"}


def matcher_cases(models):
    from models.matcher import HungarianMatcher  # reference

    m = HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)
    cases = []
    specs = [(0, 2, 100, [3, 1]), (1, 3, 100, [0, 4, 2]), (2, 2, 100, [1, 1]), (3, 4, 100, [1, 2, 3, 4]),
             (4, 2, 16, [7, 5])]
    for seed, B, Q, counts in specs:
        g = torch.Generator().manual_seed(100 + seed)
        logits = torch.randn(B, Q, 256, generator=g) * 2
        boxes = torch.cat([torch.rand(B, Q, 2, generator=g) * 0.5 + 0.25, torch.rand(B, Q, 2, generator=g) * 0.3 + 0.05], -1)
        if seed == 2:  # exact ties: every query identical -> scipy's tie rule decides
            logits[:] = logits[:, :1]
            boxes[:] = boxes[:, :1]
        targets = []
        for n in counts:
            tb = torch.cat([torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.3 + 0.05], -1)
            targets.append({"boxes": tb, "labels": torch.ones(n, dtype=torch.long)})
        T = sum(counts)
        pm = torch.zeros(T, 256)
        for r in range(T):
            lo = 1 + (r * 3) % 7
            pm[r, lo: lo + 2 + r % 3] = 1
        pm = pm / (pm.sum(-1, keepdim=True) + 1e-6)
        out = {"pred_logits": logits, "pred_boxes": boxes}
        idx = m(out, targets, pm)
        # the cost matrix itself, recomputed with the reference's own ops (matcher.py:63-82)
        from util.box_ops import box_cxcywh_to_xyxy, generalized_box_iou

        prob = logits.flatten(0, 1).softmax(-1)
        ob = boxes.flatten(0, 1)
        tb = torch.cat([t["boxes"] for t in targets])
        C = 5 * torch.cdist(ob, tb, p=1) + -(prob.unsqueeze(1) * pm.unsqueeze(0)).sum(-1) \
            + 2 * -generalized_box_iou(box_cxcywh_to_xyxy(ob), box_cxcywh_to_xyxy(tb))
        cases.append({"logits": logits, "boxes": boxes, "tgt_boxes": [t["boxes"] for t in targets], "positive_map": pm,
                      "cost": C.view(B, Q, -1), "indices": [(i.clone(), j.clone()) for i, j in idx]})
    return cases


def config1(models, tok):
    args = shims.reference_args(["--backbone", "resnet50"])
    torch.manual_seed(0)
    model, criterion, _, weight_dict = models.build_model(args)
    model.eval()
    from conftest import strided_sample
    from util.misc import NestedTensor  # reference

    images, mask, captions, targets, pm = make_batch(2, 480, 8, seed=1234, pad=True)
    # one full step of engine.py:63-88 in eval mode (no dropout): forward, criterion, weighted sum, backward
    mc = model(NestedTensor(images, mask), captions, encode_and_save=True)
    out = model(NestedTensor(images, mask), captions, encode_and_save=False, memory_cache=mc)
    losses = criterion(mc, out, targets, pm, None)
    total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)
    total.backward()
    grads = {n: p.grad.detach() for n, p in model.named_parameters() if p.grad is not None}
    with torch.no_grad():
        layers = list(out["aux_outputs"]) + [out]
        indices = [criterion.matcher(o, targets, pm) for o in layers]
    # The number format's own floor, measured on the REFERENCE: the same step under torch's CPU bf16 autocast with the
    # fp32 assignments forced (the reference calls the matcher for the main output first, then aux_0..aux_4,
    # models/mdetr.py:993,1011), per-tensor distance of its gradients from the fp32 gradients.
    order = [indices[-1]] + indices[:-1]
    calls = iter(order)
    orig_forward = criterion.matcher.forward
    criterion.matcher.forward = lambda *a, **k: next(calls)
    model.zero_grad(set_to_none=True)
    with torch.autocast("cpu", dtype=torch.bfloat16):  # the model only: the reference's criterion is not autocast-clean
        mc16 = model(NestedTensor(images, mask), captions, encode_and_save=True)
        out16 = model(NestedTensor(images, mask), captions, encode_and_save=False, memory_cache=mc16)

    def f32(d):
        return {k: (v.float() if isinstance(v, torch.Tensor) and v.is_floating_point() else
                    [f32(a) for a in v] if k == "aux_outputs" else v) for k, v in d.items()}

    out16 = f32(out16)
    losses16 = criterion(mc16, out16, targets, pm, None)
    total16 = sum(losses16[k] * weight_dict[k] for k in losses16 if k in weight_dict)
    total16.backward()
    criterion.matcher.forward = orig_forward
    from conftest import rel_err

    floor = {n: rel_err(strided_sample(p.grad.float()), strided_sample(grads[n]))
             for n, p in model.named_parameters() if p.grad is not None and n in grads and float(grads[n].norm()) > 0}
    fwd_floor = {"img_memory": rel_err(mc16["img_memory"].float(), mc["img_memory"]),
                 "pred_logits": rel_err(out16["pred_logits"].float(), out["pred_logits"]),
                 "pred_boxes": rel_err(out16["pred_boxes"].float(), out["pred_boxes"]),
                 "proj_queries": rel_err(out16["proj_queries"].float(), out["proj_queries"]),
                 "proj_tokens": rel_err(out16["proj_tokens"].float(), out["proj_tokens"])}
    print("reference under CPU bf16 autocast vs fp32, forward:", {k: f"{v:.3e}" for k, v in fwd_floor.items()})
    fl = sorted(floor.values())
    print(f"reference under CPU bf16 autocast vs fp32, gradients: median {fl[len(fl) // 2]:.3e}, max {fl[-1]:.3e}")
    sd = model.state_dict()
    small = ("contrastive_align_projection_", "class_embed.", "bbox_embed.", "query_embed.", "input_proj.bias",
             "transformer.resizer.", "transformer.decoder.norm.")
    det = lambda t: t.detach().clone()  # noqa: E731
    return {
        "batch": {"size": 480, "tokens": 8, "batch": 2, "seed": 1234, "pad": True},
        "state_checksum": {k: float(v.double().abs().sum()) for k, v in sd.items()},
        "img_memory": det(mc["img_memory"]).half(), "text_memory_resized": det(mc["text_memory_resized"]).half(),
        "mask": mc["mask"], "pos_embed_row0": mc["pos_embed"][:, 0].clone(),
        "pred_logits": torch.stack([det(o["pred_logits"]) for o in layers]),
        "pred_boxes": torch.stack([det(o["pred_boxes"]) for o in layers]),
        "proj_queries": torch.stack([det(o["proj_queries"]) for o in layers]),
        "proj_tokens": det(out["proj_tokens"]),
        "losses": {k: float(v) for k, v in losses.items()},
        "total": float(total),
        "indices": [[(i.clone(), j.clone()) for i, j in layer] for layer in indices],
        "weight_dict": dict(weight_dict),
        # reference gradients of the weighted loss sum w.r.t. every trainable tensor (436): L2 norm + an evenly
        # strided 2048-element sample of each, and the full tensor for the small head-side parameters
        "grads": {
            "norm": {k: float(g.double().norm()) for k, g in grads.items()},
            "sample": {k: strided_sample(g) for k, g in grads.items()},
            "full": {k: g.clone() for k, g in grads.items() if k.startswith(small)},
            # rel_err (same strided sample) of the reference's own gradients under torch CPU bf16 autocast
            "bf16_floor": floor,
        },
        "bf16_forward_floor": fwd_floor,
    }


def config3_small(models, tok):
    """BASELINE config 3's recipe (frozen detector + mask head, no aux / contrastive losses) at a size the CPU suite
    can afford: ResNet-50, 2 x 3 x 128 x 128 ragged, Bernoulli(.5) target masks."""
    args = shims.reference_args(["--backbone", "resnet50", "--mask_model", "smallconv", "--frozen_weights", "unused",
                                 "--no_aux_loss", "--no_contrastive_align_loss"])
    torch.manual_seed(0)
    model, criterion, _, weight_dict = models.build_model(args)
    model.eval()
    from util.misc import NestedTensor  # reference

    images, mask, captions, targets, pm = make_batch(2, 128, 8, seed=7, pad=True, masks=True)
    mc = model(NestedTensor(images, mask), captions, encode_and_save=True)
    out = model(NestedTensor(images, mask), captions, encode_and_save=False, memory_cache=mc)
    losses = criterion(mc, out, targets, pm, None)
    total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)
    total.backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    with torch.no_grad():
        indices = criterion.matcher(out, targets, pm)
    sd = model.state_dict()
    return {
        "batch": {"size": 128, "tokens": 8, "batch": 2, "seed": 7, "pad": True, "masks": True},
        "state_checksum": {k: float(v.double().abs().sum()) for k, v in sd.items()},
        "pred_masks": out["pred_masks"].detach().clone(),
        "pred_logits": out["pred_logits"].detach().clone(), "pred_boxes": out["pred_boxes"].detach().clone(),
        "losses": {k: float(v) for k, v in losses.items()},
        "indices": [(i.clone(), j.clone()) for i, j in indices],
        "grads": grads,  # mask-branch parameters only: the detector is frozen
        "weight_dict": dict(weight_dict),
    }


def config5_softkd_small(models, tok):
    """Distillation branch of SetCriterion (models/mdetr.py:887-989) with the soft-KD loss: teacher = deepcopy(student)
    at initialisation (main.py:322), different inputs; ResNet-50, 2 x 3 x 128 x 128, eval mode.  Stores the fp32
    predictions of both models (all decoder layers), the 66 loss terms, and d(sum of soft-KD terms)/d(student logits)."""
    from copy import deepcopy

    args = shims.reference_args(["--backbone", "resnet50", "--distillation", "--softkd_loss", "--softkd_coef", "50"])
    torch.manual_seed(0)
    model, criterion, _, weight_dict = models.build_model(args)
    model.eval()
    model_noun = deepcopy(model)
    from util.misc import NestedTensor  # reference

    bn = make_batch(2, 128, 8, seed=21, pad=True)
    bs = make_batch(2, 128, 8, seed=22, pad=False)
    outs, mcs = [], []
    with torch.no_grad():
        for m, (images, mask, captions, targets, pm) in ((model_noun, bn), (model, bs)):
            mc = m(NestedTensor(images, mask), captions, encode_and_save=True)
            outs.append(m(NestedTensor(images, mask), captions, encode_and_save=False, memory_cache=mc))
            mcs.append(mc)
        losses = criterion(mcs, outs, [bn[3], bs[3]], [bn[4], bs[4]], None)

    def stacked(o, k):
        return torch.stack([a[k] for a in o["aux_outputs"]] + [o[k]])

    # gradient of the soft-KD terms alone w.r.t. the student's logits of every layer
    leaf = []
    for o in outs:
        d = {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in o.items() if k != "aux_outputs"}
        d["aux_outputs"] = [{k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in a.items()}
                            for a in o["aux_outputs"]]
        leaf.append(d)
    sth_logits = [a["pred_logits"] for a in leaf[1]["aux_outputs"]] + [leaf[1]["pred_logits"]]
    for t in sth_logits:
        t.requires_grad_(True)
    l2 = criterion(mcs, leaf, [bn[3], bs[3]], [bn[4], bs[4]], None)
    kd = sum(v for k, v in l2.items() if k.startswith("loss_softkd"))
    kd.backward()
    # soft-KD in isolation on hand-made predictions where the two-class probabilities differ substantially
    from models.matcher import HungarianMatcher  # reference

    hm = HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)
    kd_cases = []
    for seed, B, Q, counts, tie in ((0, 2, 100, [2, 1], False), (1, 3, 100, [1, 4, 3], False), (2, 2, 16, [3, 2], True),
                                    (3, 2, 100, [0, 2], False)):
        g = torch.Generator().manual_seed(500 + seed)

        def preds():
            lg = torch.randn(B, Q, 256, generator=g) * 2
            lg[..., -1] += 4 + 2 * torch.randn(B, Q, generator=g)
            bx = torch.cat([torch.rand(B, Q, 2, generator=g) * 0.5 + 0.25, torch.rand(B, Q, 2, generator=g) * 0.3 + 0.05], -1)
            return lg, bx

        ln, bxn = preds()
        ls, bxs = preds()
        if tie:  # identical unmatched predictions: the LSAP tie rule decides the pairing
            ln[:, 4:] = ln[:, 4:5]
            bxn[:, 4:] = bxn[:, 4:5]
        tg = []
        for n in counts:
            tb = torch.cat([torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.3 + 0.05], -1)
            tg.append({"boxes": tb, "labels": torch.ones(n, dtype=torch.long)})
        T = sum(counts)
        pmap = torch.zeros(T, 256)
        pmap[:, 1:5] = 0.25
        on = {"pred_logits": ln, "pred_boxes": bxn}
        os_ = {"pred_logits": ls.clone().requires_grad_(True), "pred_boxes": bxs}
        idx_n, idx_s = hm(on, tg, pmap), hm({"pred_logits": ls, "pred_boxes": bxs}, tg, pmap)
        criterion.args.num_queries = Q
        val = criterion.loss_softkd([None, None], [on, os_], [tg, tg], [pmap, pmap], [idx_n, idx_s], [1.0, 1.0])["loss_softkd"]
        val.backward()
        kd_cases.append({"logits_noun": ln, "boxes_noun": bxn, "logits_sth": ls, "boxes_sth": bxs,
                         "tgt_boxes": [t["boxes"] for t in tg], "idx_noun": idx_n, "idx_sth": idx_s,
                         "loss": float(val), "grad_logits_sth": os_["pred_logits"].grad.clone()})
    criterion.args.num_queries = 100
    return {
        "softkd_cases": kd_cases,
        "batch_noun": {"size": 128, "tokens": 8, "batch": 2, "seed": 21, "pad": True},
        "batch_sth": {"size": 128, "tokens": 8, "batch": 2, "seed": 22, "pad": False},
        "noun": {k: stacked(outs[0], k) for k in ("pred_logits", "pred_boxes", "proj_queries")},
        "sth": {k: stacked(outs[1], k) for k in ("pred_logits", "pred_boxes", "proj_queries")},
        "noun_proj_tokens": outs[0]["proj_tokens"], "sth_proj_tokens": outs[1]["proj_tokens"],
        "losses": {k: float(v) for k, v in losses.items()},
        "loss_order": list(losses.keys()),
        "softkd_grad_sth_logits": torch.stack([t.grad for t in sth_logits]),
        "weight_dict": dict(weight_dict),
    }


def nsthl2_cases(models, tok):
    """SetCriterion.loss_nsthl2 (models/mdetr.py:668-781) called directly on hand-made memory caches: random
    `text_memory` [T, B, 256] of teacher and student, captions tokenised by the synthetic tokenizer, targets with
    `noun_tokens_positive` spans, assignments with the given lengths (an image without targets included).  Stores the
    loss and its gradient w.r.t. the student's text memory."""
    args = shims.reference_args(["--backbone", "resnet50", "--distillation", "--nsthl2_loss"])
    torch.manual_seed(0)
    _, criterion, _, weight_dict = models.build_model(args)
    cases = []
    for seed, counts_noun, counts_sth in ((0, [2, 1, 3], [2, 1, 3]), (1, [1, 0, 2, 4], [1, 0, 2, 4]), (2, [0, 0], [0, 0]),
                                          (3, [0, 2], [1, 2])):
        g = torch.Generator().manual_seed(900 + seed)
        B = len(counts_sth)
        caps, tgts, mcs, outs, idx = [], [], [], [], []
        for which, counts in (("noun", counts_noun), ("sth", counts_sth)):
            _, _, captions, targets, _ = make_batch(B, 32, 10, seed=70 + seed + (0 if which == "noun" else 50))
            if which == "noun":  # teacher captions name the object: a different span than the student's "something"
                captions = [c[:-9] + "hammering"[: 9] for c in captions]
            for i, t in enumerate(targets):
                n = counts[i]
                cap = captions[i]
                t["boxes"] = t["boxes"][:1].repeat(max(n, 1), 1)[:n]
                t["labels"] = torch.ones(n, dtype=torch.long)
                spans = [[[len(cap) - 9, len(cap)]], [[len(cap) - 9, len(cap) - 4]], [[len(cap) - 5, len(cap)], [0, 2]],
                         [[len(cap) - 9, len(cap) - 7]]]
                t["noun_tokens_positive"] = [spans[j % 4] for j in range(n)]
                t["tokens_positive"] = [[[0, len(cap)]] for _ in range(n)]
            tokenized = tok.batch_encode_plus(captions, padding="longest", return_tensors="pt")
            T = tokenized["input_ids"].shape[1]
            text = torch.randn(T, B, 256, generator=g)
            if which == "sth":
                text.requires_grad_(True)
            caps.append(captions)
            tgts.append(targets)
            mcs.append({"text_memory": text})
            outs.append({"proj_queries": torch.zeros(B, 100, 64), "proj_tokens": torch.zeros(B, T, 64), "tokenized": tokenized})
            idx.append([(torch.arange(n), torch.arange(n)) for n in counts])
        val = criterion.loss_nsthl2(mcs, outs, tgts, [None, None], idx, [1.0, 1.0], None)["loss_nsthl2"]
        grad = None
        if val.requires_grad:
            val.backward()
            grad = mcs[1]["text_memory"].grad.clone()
        cases.append({"captions": caps, "counts": [counts_noun, counts_sth],
                      "noun_tokens_positive": [[t["noun_tokens_positive"] for t in tg] for tg in tgts],
                      "text_noun": mcs[0]["text_memory"].detach().clone(), "text_sth": mcs[1]["text_memory"].detach().clone(),
                      "loss": float(val), "grad_text_sth": grad})
    return {"cases": cases, "weight_keys": [k for k in weight_dict if "nsthl2" in k]}


def cluster_cases(models, tok):
    """ClusterCriterion (models/mdetr.py:29-312) driven like engine.py:182-190 for a few steps on random features.
    The reference constructor needs `.cuda()` and a process group: here `.cuda()` is patched to the identity and a
    1-rank gloo group is created, so the module runs on the CPU.  memory_size = 8 so that the banks fill up within the
    run and both update branches (FIFO while filling, LSAP replacement afterwards) are exercised."""
    import os

    import numpy as np
    import torch.distributed as dist

    from models.mdetr import ClusterCriterion  # reference

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    if not dist.is_initialized():
        dist.init_process_group("gloo", rank=0, world_size=1)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        args = shims.reference_args(["--cluster", "--cluster_memory_size", "8", "--cluster_num", "3",
                                     "--train_batch_size", "4"])
        D, T, S_img, B = 32, 12, 5, 4
        torch.manual_seed(3)
        cc = ClusterCriterion(feature_dim=D, memory_size=8, cluster_num=3, task_count=14, args=args)
        init = {k: v.clone() for k, v in cc.state_dict().items()}
        np.random.seed(7)
        g = torch.Generator().manual_seed(11)
        steps = []
        nouns = ["chair", "mug", "knife", "spade"]
        for step in range(10):
            tasks = [1 + int(torch.randint(0, 3, (1,), generator=g)) for _ in range(B)]
            cap_n = [f"sit {nouns[(step + i) % 4]}" for i in range(B)]
            cap_s = ["sit something"[: T - 2] if False else "sit something"[-(T - 2):] for _ in range(B)]
            tg_n, tg_s = [], []
            for i in range(B):
                nbox = 0 if (step == 2 and i == 1) else 1 + (i % 2)
                s = cap_n[i].find(" ") + 1
                tg_n.append({"boxes": torch.rand(nbox, 4, generator=g), "dataset_name": f"tdod_{tasks[i]}",
                             "noun_tokens_positive": [[[s, len(cap_n[i])]] for _ in range(nbox)]})
                tg_s.append({"boxes": torch.rand(max(nbox, 1), 4, generator=g), "dataset_name": f"tdod_{tasks[i]}"})
            rec = {"tasks": tasks, "cap_n": cap_n, "cap_s": cap_s,
                   "nbox": [len(t["boxes"]) for t in tg_n],
                   "noun_tokens_positive": [t["noun_tokens_positive"] for t in tg_n]}
            for tag, caps, tg in (("n", cap_n, tg_n), ("s", cap_s, tg_s)):
                img = torch.randn(S_img + T, B, D, generator=g)
                mc = {"img_memory": img.clone().requires_grad_(tag == "s"), "tokenized": tok(caps)}
                Tt = mc["tokenized"]["input_ids"].shape[1]
                mc["text_memory"] = mc["img_memory"][-Tt:]
                rec["img_" + tag] = img
                if tag == "n":
                    mc = cc.update_memory(mc, tg, caps)
                    rec["mod_n"] = mc["img_memory_mod"].detach().clone()
                else:
                    mc, loss = cc(mc, tg, caps)
                    rec["mod_s"] = mc["img_memory_mod"].detach().clone()
                    rec["loss"] = {k: float(v) for k, v in loss.items()}
                    (loss["loss_cluster_feature"] * 3.0 + mc["img_memory_mod"].sum()).backward()
                    rec["grad_img_s"] = mc["img_memory"].grad.clone()
            rec["state"] = {k: v.clone() for k, v in cc.state_dict().items()}
            steps.append(rec)
        return {"init": init, "steps": steps, "dims": {"D": D, "T": T, "S_img": S_img, "B": B, "memory_size": 8,
                                                     "cluster_num": 3}, "numpy_seed": 7}
    finally:
        torch.Tensor.cuda = orig_cuda


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="comma-separated subset of: matcher,config1,config3,config5,cluster,nsthl2")
    only = set(filter(None, ap.parse_args().only.split(",")))
    torch.set_num_threads(8)
    OUT.mkdir(parents=True, exist_ok=True)
    tok = CharTokenizer()
    models = shims.load_reference(tok)
    v = versions()
    jobs = [("matcher", "matcher_cases.pt", lambda: {"cases": matcher_cases(models)}),
            ("config1", "config1_r50.pt", lambda: config1(models, tok)),
            ("config3", "config3_r50_segm_small.pt", lambda: config3_small(models, tok)),
            ("config5", "config5_softkd_small.pt", lambda: config5_softkd_small(models, tok)),
            ("cluster", "cluster_cases.pt", lambda: cluster_cases(models, tok)),
            ("nsthl2", "nsthl2_cases.pt", lambda: nsthl2_cases(models, tok))]
    for tag, fname, fn in jobs:
        if only and tag not in only:
            continue
        torch.save({"versions": v, **fn()}, OUT / fname)
    for f in sorted(OUT.glob("*.pt")):
        print(f.name, f.stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    main()
