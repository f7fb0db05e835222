"""Checkpoint compatibility with the reference's on-disk format (reference main.py:456-531, 635-653).

A reference checkpoint is `torch.save({"model", "model_ema", "model_noun", "model_noun_ema", "optimizer", "epoch", "args",
"cluster_criterion"})`; its state-dict keys are the ones `toist_b200.models` keeps, so weights trained
with the reference (A100-era runs, the published TOIST / MDETR checkpoints) load into the B200 model and vice versa.
The three loading modes of main.py are mirrored one to one, including their quirks:

  load_weights          --load            prefers `model_ema` whenever the key exists, strict=False     (main.py:456-473)
  load_frozen_weights   --frozen_weights  into `model.detr`, prefers a non-None `model_ema`             (main.py:475-489)
  resume                --resume          strips the `detr.` prefix when a segmentation checkpoint is resumed into a
                                          detection model; optimizer / epoch / EMA / cluster state      (main.py:492-531)

Keys of other library versions that have no counterpart (transformers 4.5.1 stores `embeddings.position_ids` as a
persistent buffer, torchvision's BatchNorm `num_batches_tracked`) are ignored exactly as `strict=False` ignores them in
the reference; tensors whose shape differs raise, as they do there.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Dict, Optional, Union

import torch

Checkpoint = Union[str, "os.PathLike[str]", Dict[str, Any]]


def read(checkpoint: Checkpoint) -> Dict[str, Any]:
    """A checkpoint dict, or a path to one (always mapped to the CPU first, as main.py does)."""
    if isinstance(checkpoint, dict):
        return checkpoint
    path = str(checkpoint)
    if path.startswith("https"):
        return torch.hub.load_state_dict_from_url(path, map_location="cpu", check_hash=True)
    return torch.load(path, map_location="cpu", weights_only=False)


def _without_ddp(model):
    return model.module if hasattr(model, "module") else model


def strip_detr_prefix(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """main.py:499-505: keep only the `detr.*` entries of a DETRsegm state dict, without the prefix."""
    return {k[5:]: v for k, v in state.items() if k[:5] == "detr."}


def is_segmentation_state(state: Dict[str, torch.Tensor]) -> bool:
    return "mask_head.adapter3.bias" in state  # the reference's own test (main.py:498)


def load_weights(model, checkpoint: Checkpoint):
    """`--load`: weights only, EMA weights preferred, strict=False.  Returns torch's (missing, unexpected) key lists."""
    ck = read(checkpoint)
    state = ck["model_ema"] if "model_ema" in ck else ck["model"]
    return _without_ddp(model).load_state_dict(state, strict=False)


def load_frozen_weights(model, checkpoint: Checkpoint, cluster_criterion=None):
    """`--frozen_weights`: a detection checkpoint into the detector of a DETRsegm model."""
    ck = read(checkpoint)
    m = _without_ddp(model)
    if not hasattr(m, "detr"):
        raise ValueError("--frozen_weights loads into model.detr: build the model with --mask_model smallconv")
    state = ck["model_ema"] if ck.get("model_ema") is not None else ck["model"]
    res = m.detr.load_state_dict(state, strict=False)
    if cluster_criterion is not None and "cluster_criterion" in ck:
        cluster_criterion.load_state_dict(ck["cluster_criterion"], strict=False)
    return res


def resume(model, checkpoint: Checkpoint, masks: bool, optimizer=None, model_ema=None, cluster_criterion=None,
           eval_only: bool = False, want_ema: bool = False) -> Dict[str, Any]:
    """`--resume`.  Returns {"start_epoch": epoch + 1 | None, "model_ema": the EMA model (a fresh deepcopy of `model`
    when `want_ema` and the checkpoint has none, as main.py:516-518), "missing": [...], "unexpected": [...]}."""
    ck = read(checkpoint)
    m = _without_ddp(model)
    state = ck["model"]
    if not masks and is_segmentation_state(state):
        state = strip_detr_prefix(state)
    res = m.load_state_dict(state, strict=False)
    out: Dict[str, Any] = {"start_epoch": None, "model_ema": model_ema, "missing": list(res.missing_keys),
                           "unexpected": list(res.unexpected_keys)}
    if not eval_only and optimizer is not None and "optimizer" in ck and "epoch" in ck:
        optimizer.load_state_dict(ck["optimizer"])
        out["start_epoch"] = ck["epoch"] + 1
    if cluster_criterion is not None and "cluster_criterion" in ck:
        cluster_criterion.load_state_dict(ck["cluster_criterion"], strict=False)
    if want_ema or model_ema is not None:
        if "model_ema" not in ck:
            out["model_ema"] = deepcopy(m)
        else:
            ema = model_ema if model_ema is not None else deepcopy(m)
            es = ck["model_ema"]
            if not masks and is_segmentation_state(es):
                es = strip_detr_prefix(es)
            ema.load_state_dict(es, strict=False)
            out["model_ema"] = ema
    return out


def save(path, model, optimizer=None, epoch: Optional[int] = None, args=None, model_ema=None, model_noun=None,
         model_noun_ema=None, cluster_criterion=None) -> None:
    """Writes the reference's checkpoint layout (main.py:635-653) so that either implementation can resume it."""
    torch.save({
        "model": _without_ddp(model).state_dict(),
        "model_ema": model_ema.state_dict() if model_ema is not None else None,
        "model_noun": _without_ddp(model_noun).state_dict() if model_noun is not None else None,
        "model_noun_ema": model_noun_ema.state_dict() if model_noun_ema is not None else None,
        "optimizer": optimizer.state_dict() if optimizer is not None else None,
        "epoch": epoch,
        "args": args,
        "cluster_criterion": cluster_criterion.state_dict() if cluster_criterion is not None else None,
    }, path)
