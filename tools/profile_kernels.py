"""Representative launches of the hot kernels (bench shapes: R101, B=8, 640^2 -> layer3 at 40x40, S=416, Q=100) for
`ncu --set full`.  Each shape runs `--reps` times after a warm-up; CUDA-event timings are printed as JSON lines so the
same command without ncu gives the roofline numbers.

    python tools/profile_kernels.py                     # event timings
    ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -o gpurun_out/prof_gemm \
        python tools/profile_kernels.py --reps 1
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from toist_b200 import kernels as K  # noqa: E402
from toist_b200._lib import ACT_RELU  # noqa: E402

BF = torch.bfloat16


def rn(*s):
    return torch.randn(*s, device="cuda").to(BF)


def cases():
    B = 8
    out = []
    # ---- layer3 bottleneck convolutions (66 % of the step's FLOPs)
    x256 = rn(B, 40, 40, 256)
    x1024 = rn(B, 40, 40, 1024)
    w33 = rn(256, 3, 3, 256)
    w_1024_256 = rn(256, 1, 1, 1024)
    w_256_1024 = rn(1024, 1, 1, 256)
    sh256 = torch.randn(256, device="cuda")
    sh1024 = torch.randn(1024, device="cuda")
    px = B * 40 * 40
    out.append(("conv3x3_256_256_fwd", 2.0 * px * 256 * 2304, lambda: K.conv_fwd(x256, w33, sh256, pad=1, act=ACT_RELU)))
    out.append(("conv1x1_1024_256_fwd", 2.0 * px * 256 * 1024, lambda: K.conv_fwd(x1024, w_1024_256, sh256, act=ACT_RELU)))
    out.append(("conv1x1_256_1024_fwd", 2.0 * px * 256 * 1024,
                lambda: K.conv_fwd(x256, w_256_1024, sh1024, res=x1024, act=ACT_RELU)))
    out.append(("conv3x3_256_256_dgrad", 2.0 * px * 256 * 2304, lambda: K.conv_dgrad(x256, w33, (40, 40), pad=1, mask=x256)))
    dw33 = torch.zeros(256, 3, 3, 256, device="cuda")
    out.append(("conv3x3_256_256_wgrad", 2.0 * px * 256 * 2304, lambda: K.conv_wgrad(x256, x256, dw33, pad=1)))
    dw11 = torch.zeros(1024, 1, 1, 256, device="cuda")
    out.append(("conv1x1_256_1024_wgrad", 2.0 * px * 256 * 1024, lambda: K.conv_wgrad(x1024, x256, dw11)))
    # ---- large square GEMMs: the engine's main loop without wave-quantisation / launch effects
    for n in (4096, 8192):
        xa, wb = rn(n, n), rn(n, n)
        out.append((f"gemm_{n}_cubed", 2.0 * n * n * n, (lambda xa=xa, wb=wb: K.linear_fwd(xa, wb))))
    # ---- layer1 / layer2 (large pixel counts, small K)
    x64 = rn(B, 160, 160, 64)
    w64 = rn(64, 3, 3, 64)
    sh64 = torch.randn(64, device="cuda")
    out.append(("conv3x3_64_64_160_fwd", 2.0 * B * 160 * 160 * 64 * 576, lambda: K.conv_fwd(x64, w64, sh64, pad=1, act=ACT_RELU)))
    # ---- encoder GEMMs (M = 3328) and attention core (64 problems of 416 x 416 x 32)
    S, E, H = 416, 256, 8
    xs = rn(S * B, E)
    w1, w2 = rn(2048, E), rn(E, 2048)
    b1 = torch.randn(2048, device="cuda")
    hbuf = rn(S * B, 2048)
    out.append(("ffn1_3328x256x2048", 2.0 * S * B * E * 2048, lambda: K.linear_fwd(xs, w1, b1, act=ACT_RELU)))
    out.append(("ffn2_3328x2048x256", 2.0 * S * B * E * 2048, lambda: K.linear_fwd(hbuf, w2, None, res=xs)))
    q, k, v = rn(S, B, E), rn(S, B, E), rn(S, B, E)
    km = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
    out.append(("attention_core_fwd_enc", 4.0 * S * S * 32 * B * H, lambda: K.attention_fwd(q, k, v, km, H)))
    ctx, probs = K.attention_fwd(q, k, v, km, H)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    out.append(("attention_core_bwd_enc", 10.0 * S * S * 32 * B * H,
                lambda: K.attention_bwd(ctx, q, k, v, probs, H, dq, dk, dv)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--graph", action="store_true", help="capture the inner launches into a CUDA graph")
    ap.add_argument("--inner", type=int, default=1, help="back-to-back launches per timed region (amortises the "
                                                        "host launch gap; L2 is then warm for all but the first)")
    a = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {
        "bf16_tflops": 1590.0}
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for name, flops, fn in cases():
        if a.only and a.only not in name:
            continue
        for _ in range(3 if a.reps > 1 else 1):
            fn()
        torch.cuda.synchronize()
        graph = None
        if a.graph:  # replay `inner` launches as one CUDA graph: no host launch gaps inside the timed region
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    for _ in range(a.inner):
                        fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if graph is not None:
                graph.replay()
            else:
                for _ in range(a.inner):
                    fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3 / a.inner)
        ts.sort()
        t = ts[len(ts) // 2]
        print(json.dumps({"kernel": name, "us": t * 1e6, "tflops": flops / t / 1e12,
                          "frac_of_burst_peak": flops / t / 1e12 / peaks["bf16_tflops"]}), flush=True)


if __name__ == "__main__":
    main()
