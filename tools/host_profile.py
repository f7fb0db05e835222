"""cProfile of the host side of the graph-replayed training step (which Python functions the step's issue time goes to).
    python tools/host_profile.py [--steps 20]"""
from __future__ import annotations

import argparse
import cProfile
import pstats
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--top", type=int, default=70)
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

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(a.top)


if __name__ == "__main__":
    main()
