"""One training step of the bench workload inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
from __future__ import annotations

import argparse
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
    ap.add_argument("--backbone", default="resnet101")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--tokens", type=int, default=16)
    ap.add_argument("--warm", type=int, default=2)
    a = ap.parse_args()
    torch.manual_seed(0)
    model, criterion, _, wd = build_model(make_args(a.backbone, dropout=0.1))  # the bench configuration
    model.cuda().train()
    images, mask, captions, targets, pm = make_batch(a.batch, a.size, a.tokens)
    s = NestedTensor(images.cuda(), mask.cuda())
    tg, pmd = targets_to(targets, "cuda"), pm.cuda()

    def step():
        model.zero_grad(set_to_none=True)
        mc = model(s, captions, encode_and_save=True)
        out = model(s, captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, tg, pmd, None)
        total = sum(losses[k] * wd[k] for k in losses if k in wd)
        total.backward()

    for _ in range(a.warm):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
