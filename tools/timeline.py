"""GPU timeline of the graph-replayed training step through torch.profiler (CUPTI): kernel time by name, GPU busy
fraction and the gaps between kernels.   python tools/timeline.py [--steps 3]"""
from __future__ import annotations

import argparse
import collections
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from toist_b200.models import build_model  # noqa: E402
from toist_b200.synth import make_args, make_batch, targets_to  # noqa: E402
from toist_b200.util.misc import NestedTensor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-graphs", action="store_true")
    a = ap.parse_args()
    torch.manual_seed(0)
    model, criterion, _, wd = build_model(make_args("resnet101", dropout=0.1))
    model.cuda().train()
    if not a.no_graphs:
        model.enable_cuda_graphs(True)
        criterion.enable_cuda_graphs(True)
    images, mask, captions, targets, pm = make_batch(8, 640, 16)
    s = NestedTensor(images.cuda(), mask.cuda())
    tg, pmd = targets_to(targets, "cuda"), pm.cuda()

    def step():
        model.zero_grad(set_to_none=True)
        mc = model(s, captions, encode_and_save=True)
        out = model(s, captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, tg, pmd, None)
        total = sum(losses[k] * wd[k] for k in losses if k in wd)
        total.backward()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = 0.0
    cur_end = t0
    gaps = []
    for e in evs:
        st, en = e.time_range.start, e.time_range.end
        if st > cur_end:
            gaps.append((st - cur_end, e.name))
            cur_end = st
        if en > cur_end:
            busy += en - cur_end
            cur_end = en
    span = t1 - t0
    print(f"{a.steps} steps: span {span / 1e3 / a.steps:.2f} ms/step, GPU busy {busy / 1e3 / a.steps:.2f} ms/step "
          f"({100 * busy / span:.1f} %), {len(evs) // a.steps} GPU activities/step")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in evs:
        n = e.name.split("(")[0][:90]
        agg[n][0] += 1
        agg[n][1] += e.time_range.end - e.time_range.start
    print("--- GPU time by kernel (per step)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:35]:
        print(f"{t / 1e3 / a.steps:8.3f} ms {n // a.steps:5d}x  {k}")
    gaps.sort(reverse=True)
    print("--- largest idle gaps (us) and the kernel that ended them")
    for g, n in gaps[:25]:
        print(f"{g:9.1f}  {n[:100]}")
    print(f"total idle {sum(g for g, _ in gaps) / 1e3 / a.steps:.2f} ms/step in {len(gaps) // a.steps} gaps/step")


if __name__ == "__main__":
    main()
