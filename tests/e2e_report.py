"""End-to-end parity report: toist_b200 (CUDA, bf16 tensor cores) vs the fp32 oracle on the same weights and batch.

Used by tests/test_gpu_model.py (asserts) and runnable as a script for a full table:
    python tests/e2e_report.py [--backbone resnet50 --batch 2 --size 480 --tokens 8]
"""
from __future__ import annotations

import argparse
import sys
import time
import traceback
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from conftest import rel_err  # noqa: E402
from synth import make_args, make_batch, targets_to  # noqa: E402


def weighted_total(losses, weight_dict):
    return sum(losses[k] * weight_dict[k] for k in losses.keys() if k in weight_dict)


def run_oracle(sd_cpu, backbone, batch, tokenizer, weight_dict, with_grad=True, trainable=None, forced=None,
               device="cpu", autocast_bf16=False):
    """fp32 oracle (CPU by default).  `device="cuda", autocast_bf16=True` runs the same restatement under torch's
    bf16 autocast: a second, independent bf16 implementation whose distance from the fp32 result calibrates how much
    of our own distance is the number format rather than the kernels."""
    import contextlib

    from oracle import model as O

    images, mask, captions, targets, pm = batch
    images, mask, pm = images.to(device), mask.to(device), pm.to(device)
    targets = targets_to(targets, device)
    cfg = O.Config(backbone=backbone)
    sd = {k: v.clone().to(device) for k, v in sd_cpu.items()}
    if with_grad:
        for k in trainable:
            sd[k].requires_grad_(True)
    tokd = tokenizer(captions).to(device)
    ctx = torch.enable_grad() if with_grad else torch.no_grad()
    amp = torch.autocast("cuda", dtype=torch.bfloat16) if autocast_bf16 else contextlib.nullcontext()
    with ctx, amp:
        mc = O.encode(sd, cfg, images, mask, tokd["input_ids"], tokd["attention_mask"])
        out = O.decode(sd, cfg, mc)
        losses, idx = O.criterion(cfg, out, tokd, targets, pm)
        grads = {}
        if with_grad:
            # differentiate the assignment the CUDA path used (near-tie costs flip under bf16 input noise)
            flosses, _ = O.criterion(cfg, out, tokd, targets, pm, forced_indices=forced) if forced else (losses, idx)
            total = weighted_total(flosses, weight_dict)
            total.backward()
            grads = {k: sd[k].grad for k in trainable if sd[k].grad is not None}
    return mc, out, losses, idx, grads


def run_ours(model, criterion, weight_dict, batch, with_grad=True):
    from toist_b200.util.misc import NestedTensor

    images, mask, captions, targets, pm = batch
    dev = "cuda"
    samples = NestedTensor(images.to(dev), mask.to(dev))
    tg = targets_to(targets, dev)
    ctx = torch.enable_grad() if with_grad else torch.no_grad()
    with ctx:
        mc = model.encode(samples, captions, want_features=False)
        out = model.decode(mc, want_hs=True)
        losses = criterion(mc, out, tg, pm.to(dev), None)
        grads = {}
        if with_grad:
            model.zero_grad(set_to_none=True)
            total = weighted_total(losses, weight_dict)
            total.backward()
            grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    torch.cuda.synchronize()
    return mc, out, losses, criterion.last_indices(), grads


def report(backbone="resnet50", batch=2, size=480, tokens=8, pad=True, with_grad=True, verbose=True,
           calibrate=False):
    from toist_b200.models import build_model

    args = make_args(backbone)
    torch.manual_seed(0)
    model, criterion, _, weight_dict = build_model(args)
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    model.cuda().eval()
    data = make_batch(batch, size, tokens, seed=1234, pad=pad)
    rows = []

    cal = {}

    def add(name, got, ref, calib=None):
        e = rel_err(got, ref)
        rows.append((name, e))
        if verbose:
            extra = f"   (torch bf16 autocast vs fp32: {rel_err(calib, ref):.3e})" if calib is not None else ""
            print(f"  {name:58s} rel_err = {e:.3e}{extra}", flush=True)

    t0 = time.time()
    mc, out, losses, idx, grads = run_ours(model, criterion, weight_dict, data, with_grad)
    t1 = time.time()
    forced = idx[-1:] + idx[:-1]  # ours: [layer 0 .. L-1] -> oracle order [main, aux_0 ..]
    omc, oout, olosses, oidx, ograds = run_oracle(sd_cpu, backbone, data, model.transformer.tokenizer, weight_dict,
                                                  with_grad, trainable, forced)
    t2 = time.time()
    if verbose:
        print(f"ours {t1 - t0:.2f}s  oracle {t2 - t1:.2f}s")
    cmc = cout = None
    cgrads = {}
    if calibrate:
        torch.backends.cudnn.allow_tf32 = False
        cmc, cout, _, _, cgrads = run_oracle(sd_cpu, backbone, data, model.transformer.tokenizer, weight_dict,
                                             with_grad, trainable, forced, device="cuda", autocast_bf16=True)
    cget = (lambda d, k: d[k].float()) if calibrate else (lambda d, k: None)
    add("text_memory_resized", mc["text_memory_resized"], omc["text_memory_resized"], cget(cmc, "text_memory_resized"))
    add("img_memory", mc["img_memory"], omc["img_memory"], cget(cmc, "img_memory"))
    add("pos_embed", mc["pos_embed"], omc["pos_embed"])
    assert torch.equal(mc["mask"].cpu(), omc["mask"]), "key padding mask differs"
    L, B, Q = oout["hs"].shape[:3]
    hs = out["_b200_hs"].float().view(L, Q, B, -1).transpose(1, 2)
    add("hs", hs, oout["hs"])
    st = out["_b200_stacked"]
    olayers = list(oout["aux_outputs"]) + [oout]
    clayers = (list(cout["aux_outputs"]) + [cout]) if calibrate else None
    for k in ("pred_logits", "pred_boxes", "proj_queries"):
        add(k + " (all layers)", st[k], torch.stack([o[k] for o in olayers]),
            torch.stack([o[k].float() for o in clayers]) if calibrate else None)
    add("proj_tokens", st["proj_tokens"], oout["proj_tokens"], cget(cout, "proj_tokens"))
    for k in olosses:
        rows.append(("loss:" + k, abs(float(losses[k]) - float(olosses[k])) / max(1.0, abs(float(olosses[k])))))
        if verbose:
            print(f"  loss {k:40s} ours {float(losses[k]):.6f} oracle {float(olosses[k]):.6f}")
    # oracle returns [main, aux_0..]; ours [layer0 .. layer L-1]
    oidx_by_layer = oidx[1:] + oidx[:1]
    same = 0
    total = 0
    for l in range(len(idx)):
        for (r0, c0), (r1, c1) in zip(idx[l], oidx_by_layer[l]):
            total += 1
            same += int(r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist())
    rows.append(("matcher_index_agreement", same / max(total, 1)))
    if verbose:
        print(f"  matcher assignments identical on {same}/{total} (layer, image) problems (e2e; inputs differ by bf16)")
    if with_grad:
        keys = sorted(ograds)
        worst = []
        for k in keys:
            if k not in grads:
                rows.append(("grad:" + k, float("inf")))
                print("  MISSING grad", k)
                continue
            e = rel_err(grads[k], ograds[k])
            rows.append(("grad:" + k, e))
            worst.append((e, k))
            if k in cgrads:
                cal[k] = rel_err(cgrads[k].float(), ograds[k])
                rows.append(("cal:" + k, cal[k]))
        extra = sorted(set(grads) - set(ograds))
        if extra and verbose:
            print("  grads present only in ours:", extra[:10])
        worst.sort(reverse=True)
        if verbose:
            print("  worst gradient errors:")
            for e, k in worst[:25]:
                c = f"   (torch bf16: {cal[k]:.3e})" if k in cal else ""
                print(f"    {k:70s} {e:.3e}{c}")
            if cal:
                ratios = sorted(e / max(cal[k], 1e-12) for e, k in worst if k in cal and "key.bias" not in k)
                print(f"  ours / torch-bf16 gradient error ratio: median {ratios[len(ratios) // 2]:.2f}, "
                      f"max {ratios[-1]:.2f} over {len(ratios)} tensors")
                cos = []
                for k in ograds:
                    if k in grads and "key.bias" not in k:
                        a, b = grads[k].double().flatten().cpu(), ograds[k].double().flatten()
                        cos.append((float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)), k))
                cos.sort()
                print("  lowest gradient cosine similarities vs fp32 oracle:", [(round(c, 3), k) for c, k in cos[:5]])
            import statistics

            print(f"  median grad rel_err {statistics.median([e for e, _ in worst]):.3e} over {len(worst)} tensors")
            for tag in ("backbone.0.body.layer2.0.conv1.weight", "backbone.0.body.layer4.2.conv3.weight",
                        "transformer.text_encoder.embeddings.word_embeddings.weight",
                        "transformer.text_encoder.encoder.layer.0.attention.self.query.weight",
                        "transformer.encoder.layers.0.self_attn.in_proj_weight",
                        "transformer.decoder.layers.5.linear2.weight", "query_embed.weight", "input_proj.weight",
                        "class_embed.weight", "bbox_embed.layers.2.weight"):
                for e, k in worst:
                    if k == tag:
                        c = f"   (torch bf16: {cal[k]:.3e})" if k in cal else ""
                        print(f"    [{tag}] {e:.3e}{c}")
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--tokens", type=int, default=8)
    ap.add_argument("--no-grad", action="store_true")
    ap.add_argument("--calibrate", action="store_true", help="also run the oracle under torch bf16 autocast")
    a = ap.parse_args()
    try:
        report(a.backbone, a.batch, a.size, a.tokens, with_grad=not a.no_grad, calibrate=a.calibrate)
    except Exception:
        traceback.print_exc()
        sys.exit(1)
