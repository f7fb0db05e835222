"""PostProcess / PostProcessSegm behind the reference's interface (reference models/postprocessors.py:15-117): the
evaluation side of the path (engine.py:305-309), on the kernels of csrc/io.cu.  No CPU path: CPU tensors raise."""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn

from .. import kernels as K


class PostProcess(nn.Module):
    """Converts the model's output into the format expected by the COCO api: per image `scores` (probability of not
    being the no-object token), `labels` (all 1) and `boxes` (absolute xyxy)."""

    @torch.no_grad()
    def forward(self, outputs, target_sizes):
        out_logits, out_bbox = outputs["pred_logits"], outputs["pred_boxes"]
        assert len(out_logits) == len(target_sizes)
        assert target_sizes.shape[1] == 2
        if not out_logits.is_cuda:
            raise RuntimeError("toist_b200 PostProcess expects CUDA predictions; there is no CPU path")
        ts = target_sizes.to(out_logits.device)
        scores, labels, boxes, refexp = K.postprocess_boxes(out_logits, out_bbox, ts, outputs.get("pred_isfinal"))
        results = [{"scores": s, "labels": l, "boxes": b} for s, l, b in zip(scores, labels, boxes)]
        if refexp is not None:
            for i in range(len(results)):
                results[i]["scores_refexp"] = refexp[i]
        return results


class PostProcessSegm(nn.Module):
    """Binarised masks at the original image size, called after PostProcess.  The reference interpolates twice (mask ->
    padded batch size -> crop -> original size, models/postprocessors.py:79-107); here both bilinear passes, the sigmoid
    and the threshold are one kernel per image reading only the low-resolution mask, and only the boolean result
    crosses to the host (the reference materialises two fp32 upsampled tensors per batch first)."""

    def __init__(self, threshold=0.5):
        super().__init__()
        self.threshold = threshold

    @torch.no_grad()
    def forward(self, results, outputs, orig_target_sizes, max_target_sizes):
        assert len(orig_target_sizes) == len(max_target_sizes)
        max_h, max_w = max_target_sizes.max(0)[0].tolist()
        pm = outputs["pred_masks"]
        if pm.dim() == 5:
            pm = pm.squeeze(2)
        if not pm.is_cuda:
            raise RuntimeError("toist_b200 PostProcessSegm expects CUDA predictions; there is no CPU path")
        pm = pm.float().contiguous()
        sizes = max_target_sizes.tolist()
        origs = orig_target_sizes.tolist()
        same = all(s == sizes[0] for s in sizes) and all(o == origs[0] for o in origs) and tuple(sizes[0]) == (max_h, max_w)
        for i in range(pm.shape[0]):
            # all sizes equal: the reference skips the crop (postprocessors.py:88-95), i.e. crop == the whole stage-1 map
            crop = (max_h, max_w) if same else tuple(sizes[i])
            m = K.postprocess_masks(pm[i], (max_h, max_w), crop, tuple(origs[i]), float(self.threshold))
            results[i]["masks"] = m.unsqueeze(1).cpu()
        return results


def build_postprocessors(args, dataset_name) -> Dict[str, nn.Module]:
    postprocessors: Dict[str, nn.Module] = {"bbox": PostProcess()}
    if args.masks:
        postprocessors["segm"] = PostProcessSegm()
    return postprocessors
