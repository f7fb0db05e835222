"""The two ends of the path on the GPU (csrc/io.cu): batch assembly against the reference's
ToTensor + Normalize + NestedTensor.from_tensor_list arithmetic (bit-exact), PostProcess / PostProcessSegm against a
plain-torch restatement of reference models/postprocessors.py:15-109 (boxes bit-exact, scores 1e-6, masks >= 99.99 % of
the pixels: a bilinear value within one fp32 ulp of the threshold may fall on the other side)."""
from __future__ import annotations

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_uint8_batch_assembly_is_bit_exact():
    from toist_b200.util.misc import NestedTensor

    g = torch.Generator().manual_seed(0)
    sizes = [(37, 53), (64, 40), (1, 1), (50, 64)]
    imgs = [torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8) for h, w in sizes]
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    nt = NestedTensor.from_uint8_list([t.to(DEV) for t in imgs], mean, std)
    nt_host = NestedTensor.from_uint8_list(imgs, mean, std)  # host tensors: staged through pinned memory
    # reference: torchvision to_tensor (x / 255) then normalize (sub mean, div std), then util/misc.py:185-209
    ref = []
    for t in imgs:
        x = t.permute(2, 0, 1).float().div(255)
        x = x.sub(torch.tensor(mean)[:, None, None]).div(torch.tensor(std)[:, None, None])
        ref.append(x)
    H, W = max(h for h, _ in sizes), max(w for _, w in sizes)
    want = torch.zeros(len(imgs), 3, H, W)
    wmask = torch.ones(len(imgs), H, W, dtype=torch.bool)
    for i, x in enumerate(ref):
        want[i, :, : x.shape[1], : x.shape[2]] = x
        wmask[i, : x.shape[1], : x.shape[2]] = False
    for got in (nt, nt_host):
        assert got.tensors.shape == want.shape and got.mask.dtype == torch.bool
        assert torch.equal(got.tensors.cpu(), want) and torch.equal(got.mask.cpu(), wmask)
    r = NestedTensor.from_uint8_list([imgs[0].to(DEV)], mean, std, do_round=True)
    assert r.tensors.shape == (1, 3, 128, 128) and bool(r.mask[0, 37:].all()) and bool(r.mask[0, :, 53:].all())


def test_from_tensor_list_on_the_gpu_matches_the_host_path():
    from toist_b200.util.misc import NestedTensor

    g = torch.Generator().manual_seed(1)
    imgs = [torch.randn(3, h, w, generator=g) for h, w in ((40, 33), (17, 64), (64, 64))]
    host = NestedTensor.from_tensor_list(imgs)
    dev = NestedTensor.from_tensor_list([t.to(DEV) for t in imgs])
    assert torch.equal(dev.tensors.cpu(), host.tensors) and torch.equal(dev.mask.cpu(), host.mask)
    devr = NestedTensor.from_tensor_list([t.to(DEV) for t in imgs], do_round=True)
    assert devr.tensors.shape == (3, 3, 128, 128)


def _ref_postprocess(logits, boxes, sizes):
    prob = F.softmax(logits, -1)
    scores = 1 - prob[:, :, -1]
    cx, cy, w, h = boxes.unbind(-1)
    b = torch.stack([(cx - 0.5 * w), (cy - 0.5 * h), (cx + 0.5 * w), (cy + 0.5 * h)], dim=-1)
    img_h, img_w = sizes.unbind(1)
    return scores, b * torch.stack([img_w, img_h, img_w, img_h], dim=1)[:, None, :]


@pytest.mark.parametrize("size_dtype", [torch.int64, torch.float32])
def test_postprocess_boxes(size_dtype):
    from toist_b200.models.postprocessors import PostProcess

    g = torch.Generator().manual_seed(2)
    B, Q, Cc = 3, 100, 256
    logits = torch.randn(B, Q, Cc, generator=g) * 3
    boxes = torch.rand(B, Q, 4, generator=g)
    sizes = torch.tensor([[480, 640], [333, 500], [1333, 800]]).to(size_dtype)
    fin = torch.randn(B, Q, 1, generator=g)
    out = {"pred_logits": logits.to(DEV), "pred_boxes": boxes.to(DEV)}
    res = PostProcess()(out, sizes.to(DEV))
    s_ref, b_ref = _ref_postprocess(logits, boxes, sizes)
    assert len(res) == B and set(res[0]) == {"scores", "labels", "boxes"}
    for i in range(B):
        assert res[i]["labels"].dtype == torch.int64 and bool((res[i]["labels"] == 1).all())
        assert float((res[i]["scores"].cpu() - s_ref[i]).abs().max()) < 1e-6
        assert torch.equal(res[i]["boxes"].cpu(), b_ref[i].float())
    out["pred_isfinal"] = fin.to(DEV)
    res = PostProcess()(out, sizes.to(DEV))
    for i in range(B):
        want = s_ref[i] * fin[i].sigmoid().view(-1)
        assert float((res[i]["scores_refexp"].cpu() - want).abs().max()) < 1e-6


def _ref_segm(pred, orig, sizes, thr):
    """reference models/postprocessors.py:79-107, verbatim arithmetic in plain torch (run on the GPU by the caller)."""
    max_h, max_w = sizes.max(0)[0].tolist()
    m = F.interpolate(pred, size=(max_h, max_w), mode="bilinear", align_corners=False)
    min_h, min_w = sizes.min(0)[0].tolist()
    mo_h, mo_w = orig.min(0)[0].tolist()
    xo_h, xo_w = orig.max(0)[0].tolist()
    if min_h == max_h and min_w == max_w and mo_h == xo_h and mo_w == xo_w:
        full = F.interpolate(m, size=(mo_h, mo_w), mode="bilinear").sigmoid() > thr
        return [c.unsqueeze(1) for c in full]
    res = []
    for cur, t, tt in zip(m, sizes, orig):
        c = cur[:, : t[0], : t[1]].unsqueeze(1)
        res.append(F.interpolate(c.float(), size=tuple(tt.tolist()), mode="bilinear").sigmoid() > thr)
    return res


@pytest.mark.parametrize("uniform", [True, False])
def test_postprocess_masks(uniform):
    from toist_b200.models.postprocessors import PostProcessSegm

    g = torch.Generator().manual_seed(3)
    B, Q, h, w = 3, 20, 40, 36
    pred = (torch.randn(B, Q, h, w, generator=g) * 2).to(DEV)
    if uniform:
        sizes = torch.tensor([[160, 144]] * B)
        orig = torch.tensor([[480, 431]] * B)
    else:
        sizes = torch.tensor([[160, 144], [120, 100], [97, 144]])
        orig = torch.tensor([[480, 431], [333, 278], [200, 301]])
    results = [{} for _ in range(B)]
    got = PostProcessSegm(0.5)(results, {"pred_masks": pred.unsqueeze(2)}, orig.to(DEV), sizes.to(DEV))
    want = _ref_segm(pred, orig, sizes, 0.5)
    for i in range(B):
        m = got[i]["masks"]
        assert m.device.type == "cpu" and m.dtype == torch.bool and m.shape == want[i].shape == (Q, 1, *orig[i].tolist())
        agree = float((m == want[i].cpu()).float().mean())
        assert agree >= 0.9999, (i, agree)
        assert 0.2 < float(m.float().mean()) < 0.8


def test_eval_forward_and_postprocessors_end_to_end():
    """engine.py:253-309: eval-mode forward under no_grad + both post-processors on a small DETRsegm."""
    from toist_b200.models import build_model
    from toist_b200.models.postprocessors import build_postprocessors
    from toist_b200.synth import make_args, make_batch
    from toist_b200.util.misc import NestedTensor

    args = make_args("resnet50", masks=True, mask_model="smallconv", aux_loss=False, contrastive_align_loss=False)
    torch.manual_seed(0)
    model = build_model(args)[0].cuda().eval()
    post = build_postprocessors(args, "tdod")
    assert set(post) == {"bbox", "segm"}
    images, mask, captions, targets, _ = make_batch(2, 128, 8, seed=3, pad=True)
    with torch.no_grad():
        s = NestedTensor(images.cuda(), mask.cuda())
        mc = model(s, captions, encode_and_save=True)
        out = model(s, captions, encode_and_save=False, memory_cache=mc)
        orig = torch.tensor([[256, 256], [200, 240]], device=DEV)
        sizes = torch.tensor([[128, 128], [96, 112]], device=DEV)
        res = post["bbox"](out, orig)
        res = post["segm"](res, out, orig, sizes)
    assert res[0]["masks"].shape == (100, 1, 256, 256) and res[1]["masks"].shape == (100, 1, 200, 240)
    assert res[0]["boxes"].shape == (100, 4) and bool(torch.isfinite(res[0]["scores"]).all())
    s_ref, b_ref = _ref_postprocess(out["pred_logits"].cpu(), out["pred_boxes"].cpu(), orig.cpu())
    assert torch.equal(res[1]["boxes"].cpu(), b_ref[1].float())
    want = _ref_segm(out["pred_masks"], orig.cpu(), sizes.cpu(), 0.5)
    assert float((res[1]["masks"] == want[1].cpu()).float().mean()) >= 0.9999
