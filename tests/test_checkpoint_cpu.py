"""Checkpoint compatibility: the three loading modes of reference main.py:456-531 against
checkpoints in the reference's on-disk layout.  The first test builds the checkpoint from the UNMODIFIED reference's own
modules (this container only); the others run anywhere."""
from __future__ import annotations

from copy import deepcopy

import pytest
import torch

from oracle import shims
from toist_b200.synth import make_args
from toist_b200.util import checkpoint as ckpt


def _ours(masks: bool = False, frozen: bool = False):
    from toist_b200.models import build_model

    over = dict(masks=True, mask_model="smallconv", frozen_weights="x" if frozen else None) if masks else {}
    torch.manual_seed(123)
    return build_model(make_args("resnet50", device="cpu", **over))[0]


def _same(a, b):
    return a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)


@pytest.mark.reference
@pytest.mark.skipif(not shims.reference_available(), reason="/root/reference not present")
def test_reference_checkpoints_load(tmp_path):
    """A checkpoint written the way reference main.py:641-653 writes it, from the reference's own segmentation model:
    --load takes the EMA weights, --resume into a detection model strips `detr.`, --frozen_weights fills model.detr."""
    from toist_b200.tokenizer import CharTokenizer

    models = shims.load_reference(CharTokenizer())
    torch.manual_seed(0)
    ref_det = models.build_model(shims.reference_args(["--backbone", "resnet50"]))[0]
    torch.manual_seed(1)
    ref_seg = models.build_model(shims.reference_args(["--backbone", "resnet50", "--mask_model", "smallconv"]))[0]
    ema = deepcopy(ref_det)
    with torch.no_grad():
        for p in ema.parameters():
            p.mul_(0.5)
    # transformers 4.5.1 checkpoints also carry `position_ids` and torchvision BN `num_batches_tracked`: add them back
    det_state = dict(ref_det.state_dict())
    det_state["transformer.text_encoder.embeddings.position_ids"] = torch.arange(514).unsqueeze(0)
    det_state["backbone.0.body.bn1.num_batches_tracked"] = torch.tensor(7)
    opt = torch.optim.AdamW(ref_det.parameters(), lr=1e-4)
    f_det, f_seg = tmp_path / "det.pth", tmp_path / "seg.pth"
    torch.save({"model": det_state, "model_ema": ema.state_dict(), "model_noun": None, "model_noun_ema": None,
                "optimizer": opt.state_dict(), "epoch": 4, "args": None, "cluster_criterion": None}, f_det)
    torch.save({"model": ref_seg.state_dict(), "model_ema": None, "optimizer": opt.state_dict(), "epoch": 9}, f_seg)

    ours = _ours()
    res = ckpt.load_weights(ours, str(f_det))                         # --load: EMA preferred (main.py:458-459)
    assert not res.missing_keys and not res.unexpected_keys
    assert _same(dict(ours.state_dict()), dict(ema.state_dict()))
    ours2 = _ours()
    info = ckpt.resume(ours2, str(f_det), masks=False, want_ema=True)  # --resume, detection into detection
    assert sorted(info["unexpected"]) == ["transformer.text_encoder.embeddings.position_ids"] and not info["missing"]
    assert _same(dict(ours2.state_dict()), dict(ref_det.state_dict()))
    assert _same(dict(info["model_ema"].state_dict()), dict(ema.state_dict()))
    ours3 = _ours()
    info = ckpt.resume(ours3, str(f_seg), masks=False)                 # segmentation checkpoint into a detector
    assert not info["missing"] and not info["unexpected"]
    assert _same(dict(ours3.state_dict()), dict(ref_seg.detr.state_dict()))
    seg = _ours(masks=True, frozen=True)
    res = ckpt.load_frozen_weights(seg, str(f_det))                     # --frozen_weights: EMA is not None -> EMA
    assert _same(dict(seg.detr.state_dict()), dict(ema.state_dict()))
    seg2 = _ours(masks=True)
    info = ckpt.resume(seg2, str(f_seg), masks=True)                    # segmentation into segmentation: as is
    assert not info["missing"] and _same(dict(seg2.state_dict()), dict(ref_seg.state_dict()))


def test_round_trip_and_resume_state(tmp_path):
    """save() writes the reference layout; resume restores weights, EMA, optimizer state and the epoch counter, and
    skips the optimizer under --eval (main.py:509-511)."""
    m = _ours()
    ema = deepcopy(m)
    with torch.no_grad():
        for p in ema.parameters():
            p.add_(1.0)
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad][:4], lr=3e-4)
    for p in opt.param_groups[0]["params"]:
        p.grad = torch.ones_like(p)
    opt.step()
    f = tmp_path / "checkpoint.pth"
    ckpt.save(f, m, optimizer=opt, epoch=11, model_ema=ema)
    raw = torch.load(f, map_location="cpu", weights_only=False)
    assert set(raw) == {"model", "model_ema", "model_noun", "model_noun_ema", "optimizer", "epoch", "args", "cluster_criterion"}
    m2 = _ours()
    with torch.no_grad():
        for p in m2.parameters():
            p.zero_()
    opt2 = torch.optim.AdamW([p for p in m2.parameters() if p.requires_grad][:4], lr=1.0)
    info = ckpt.resume(m2, str(f), masks=False, optimizer=opt2, want_ema=True)
    assert info["start_epoch"] == 12 and opt2.param_groups[0]["lr"] == 3e-4
    assert _same(dict(m2.state_dict()), dict(m.state_dict()))
    assert _same(dict(info["model_ema"].state_dict()), dict(ema.state_dict()))
    info = ckpt.resume(_ours(), str(f), masks=False, optimizer=torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))]),
                       eval_only=True)
    assert info["start_epoch"] is None
    # a checkpoint without EMA weights: the EMA model restarts from the loaded model (main.py:516-518)
    raw.pop("model_ema")
    m3 = _ours()
    info = ckpt.resume(m3, raw, masks=False, want_ema=True)
    assert _same(dict(info["model_ema"].state_dict()), dict(m3.state_dict())) and info["model_ema"] is not m3
    # --load without `model_ema` falls back to `model`
    m4 = _ours()
    ckpt.load_weights(m4, {"model": m.state_dict()})
    assert _same(dict(m4.state_dict()), dict(m.state_dict()))
    # shape mismatches are errors, as in the reference
    bad = {k: v for k, v in m.state_dict().items()}
    bad["class_embed.bias"] = torch.zeros(3)
    with pytest.raises(RuntimeError):
        ckpt.load_weights(_ours(), {"model": bad})
