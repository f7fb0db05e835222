"""Per-CTA phase timeline of toist::gemm_kernel for the bench's layer3 shapes (toist_debug_gemm_trace: clock64 stamps).

    python tools/gemm_trace.py [--only conv3x3] [--chain 6]      (TOIST_GEMM_2SM=1 / TOIST_GEMM_MIN_CTAS=96 select variants)
Prints, per shape: the median over CTAs of each phase's duration in SM clocks; with --chain N the same launch is captured
N times back to back in a CUDA graph (each launch with its own trace region) and the global-timer view of the launch
boundaries is printed: when the first / last CTA of launch k+1 entered relative to the last exit of launch k."""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

from toist_b200 import _lib  # noqa: E402

NAMES = ["setup", "dep wait", "first operands", "main loop issue", "drain to accum", "epilogue tiles", "store"]
REGION = 16 * 4096  # int64 slots per launch


PNAMES = ["setup", "dep wait", "first operands", "all main loops", "(first accumulator)", "epilogue end after last MMA"]


def phases(t, label):
    lead = t[t[:, 4] != 0]  # CTAs that issued MMAs (cta_group::2: the leaders)
    print(f"   {label}: {t.shape[0]} CTAs traced ({lead.shape[0]} issuing MMAs)")
    if 0 < int(lead[:, 7].max()) < 100000:  # persistent kernel: slot 7 is the CTA's tile count
        tiles = lead[:, 7].float()
        print(f"      persistent kernel: tiles per CTA min {int(tiles.min())} median {int(tiles.median())} max {int(tiles.max())}"
              f" (sum {int(tiles.sum())})")
        for nm, a, b in (("setup", 0, 1), ("dep wait", 1, 2), ("first operands", 2, 3), ("MMA warp: all tiles", 3, 4),
                         ("first accumulator", 2, 5), ("epilogue: all tiles", 5, 6), ("epilogue after last MMA", 4, 6),
                         ("CTA lifetime", 0, 6)):
            col = (lead[:, b] - lead[:, a]).float()
            print(f"      {nm:24s} median {col.median():9.0f}  min {col.min():9.0f}  max {col.max():9.0f} clk")
        per = ((lead[:, 6] - lead[:, 5]).float() / tiles)
        print(f"      {'epilogue clocks per tile':24s} median {per.median():9.0f}  min {per.min():9.0f}  max {per.max():9.0f}")
        return
    d = lead[:, 1:8] - lead[:, 0:7]
    for i, nm in enumerate(NAMES):
        col = d[:, i].float()
        print(f"      {nm:18s} median {col.median():9.0f}  min {col.min():9.0f}  max {col.max():9.0f} clk")
    tot = (lead[:, 7] - lead[:, 0]).float()
    print(f"      {'CTA lifetime':18s} median {tot.median():9.0f}  min {tot.min():9.0f}  max {tot.max():9.0f} clk")


def main():
    from profile_kernels import cases

    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="conv")
    ap.add_argument("--chain", type=int, default=0)
    a = ap.parse_args()
    L = _lib.load()
    n_regions = max(1, a.chain)
    buf = torch.zeros(n_regions * REGION, dtype=torch.int64, device="cuda")
    for name, flops, fn in cases():
        if a.only not in name or "wgrad" in name:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        buf.zero_()
        print(f"== {name}")
        if not a.chain:
            L.toist_debug_gemm_trace(buf.data_ptr())
            fn()
            torch.cuda.synchronize()
            L.toist_debug_gemm_trace(None)
            t = buf.view(-1, 16).cpu()
            phases(t[t[:, 0] != 0], "single launch")
            continue
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for k in range(a.chain):
                    L.toist_debug_gemm_trace(buf.data_ptr() + 8 * k * REGION)
                    fn()
            L.toist_debug_gemm_trace(None)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        t = buf.view(a.chain, -1, 16).cpu()
        prev_exit = None
        for k in range(a.chain):
            tk = t[k][t[k][:, 0] != 0]
            ent, ex = tk[:, 8], tk[:, 9]
            line = (f"   launch {k}: {tk.shape[0]} CTAs; entries span {int(ent.max() - ent.min())} ns; first entry -> last "
                    f"exit {int(ex.max() - ent.min())} ns; CTA lifetime median {int((ex - ent).median())} ns")
            if prev_exit is not None:
                line += (f"; first entry {int(ent.min() - prev_exit):+d} ns / last entry {int(ent.max() - prev_exit):+d} ns "
                         f"vs previous launch's last exit; period {int(ex.max() - prev_exit)} ns")
            print(line)
            prev_exit = int(ex.max())
        phases(t[a.chain // 2][t[a.chain // 2][:, 0] != 0], f"launch {a.chain // 2} of the chain")


if __name__ == "__main__":
    main()
