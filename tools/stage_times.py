"""Per-stage device time of the graph-replayed training step (CUDA events around every graph replay) plus the host
wall time of the step: shows which stage the step spends its time in and how much of the step the host is late.
    python tools/stage_times.py [--steps 5]"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from toist_b200 import runtime as R  # noqa: E402
from toist_b200.models import build_model  # noqa: E402
from toist_b200.synth import make_args, make_batch, targets_to  # noqa: E402
from toist_b200.util.misc import NestedTensor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    torch.manual_seed(0)
    model, criterion, _, wd = build_model(make_args("resnet101", dropout=0.1))
    model.cuda().train()
    model.enable_cuda_graphs(True)
    criterion.enable_cuda_graphs(True)
    criterion.enable_fused_loss_sum(True)
    import os

    from toist_b200.util.optim import FusedAdamW

    if os.environ.get("TOIST_DIRECT", "1") != "0":
        model.enable_direct_grads(True)
    optimizer = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    images, mask, captions, targets, pm = make_batch(8, 640, 16)
    s = NestedTensor(images.cuda(), mask.cuda())
    tg, pmd = targets_to(targets, "cuda"), pm.cuda()

    def step():
        optimizer.zero_grad()
        mc = model(s, captions, encode_and_save=True)
        out = model(s, captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, tg, pmd, None)
        total = sum(losses[k] * wd[k] for k in losses if k in wd)
        total.backward()

    def traced_step(log):
        """The same step with (label, host clock, CUDA event) probes between the calls: shows whether a gap on the
        device is the host being late or device-side work / dependencies."""
        def probe(label):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            log.append((label, time.perf_counter(), ev))

        probe("step start")
        optimizer.zero_grad()
        probe("zero_grad done")
        mc = model(s, captions, encode_and_save=True)
        probe("phase A issued")
        out = model(s, captions, encode_and_save=False, memory_cache=mc)
        probe("phase B issued")
        losses = criterion(mc, out, tg, pmd, None)
        probe("criterion issued")
        total = sum(losses[k] * wd[k] for k in losses if k in wd)
        probe("weighted sum issued")
        total.backward()
        probe("backward issued")

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    R.GraphCache.timing = []
    marks = []
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append(e)
        step()
    host = (time.perf_counter() - t0) / a.steps
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append(e)
    torch.cuda.synchronize()
    print(f"host issue time {host * 1e3:.2f} ms/step; device span {marks[0].elapsed_time(marks[-1]) / a.steps:.2f} ms/step")
    rows = R.GraphCache.timing
    per = len(rows) // a.steps
    print(f"{per} graph replays per step")
    last = rows[-per:]
    base = marks[-2]
    for ph, sig, n, e0, e1 in last:
        print(f"  start {base.elapsed_time(e0):7.3f} ms  dur {e0.elapsed_time(e1):7.3f} ms  {n:4d} launches  {ph} {sig}")
    # host clock vs device clock at the call boundaries of a step, once with the host running ahead (back-to-back steps)
    R.GraphCache.timing = None
    for _ in range(3):
        step()
    log = []
    traced_step(log)
    torch.cuda.synchronize()
    h0, e0 = log[0][1], log[0][2]
    print("probe                      host ms   device ms   (device time = when the GPU reached the probe)")
    for label, h, ev in log:
        print(f"  {label:24s} {(h - h0) * 1e3:7.3f}   {e0.elapsed_time(ev):8.3f}")


if __name__ == "__main__":
    main()
