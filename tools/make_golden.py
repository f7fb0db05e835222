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
    from util.misc import NestedTensor  # reference

    images, mask, captions, targets, pm = make_batch(2, 480, 8, seed=1234, pad=True)
    with torch.no_grad():
        mc = model(NestedTensor(images, mask), captions, encode_and_save=True)
        out = model(NestedTensor(images, mask), captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, targets, pm, None)
        layers = list(out["aux_outputs"]) + [out]
        indices = [criterion.matcher(o, targets, pm) for o in layers]
    sd = model.state_dict()
    return {
        "batch": {"size": 480, "tokens": 8, "batch": 2, "seed": 1234, "pad": True},
        "state_checksum": {k: float(v.double().abs().sum()) for k, v in sd.items()},
        "img_memory": mc["img_memory"].half(), "text_memory_resized": mc["text_memory_resized"].half(),
        "mask": mc["mask"], "pos_embed_row0": mc["pos_embed"][:, 0].clone(),
        "pred_logits": torch.stack([o["pred_logits"] for o in layers]),
        "pred_boxes": torch.stack([o["pred_boxes"] for o in layers]),
        "proj_queries": torch.stack([o["proj_queries"] for o in layers]),
        "proj_tokens": out["proj_tokens"],
        "losses": {k: float(v) for k, v in losses.items()},
        "indices": [[(i.clone(), j.clone()) for i, j in layer] for layer in indices],
        "weight_dict": dict(weight_dict),
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


def main():
    torch.set_num_threads(8)
    OUT.mkdir(parents=True, exist_ok=True)
    tok = CharTokenizer()
    models = shims.load_reference(tok)
    v = versions()
    torch.save({"versions": v, "cases": matcher_cases(models)}, OUT / "matcher_cases.pt")
    torch.save({"versions": v, **config1(models, tok)}, OUT / "config1_r50.pt")
    torch.save({"versions": v, **config3_small(models, tok)}, OUT / "config3_r50_segm_small.pt")
    for f in sorted(OUT.glob("*.pt")):
        print(f.name, f.stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    main()
