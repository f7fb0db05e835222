"""GPU self-test of the implicit-GEMM engine: each case runs in its own subprocess (a trapping kernel kills the CUDA
context), compares against torch fp32 math on the same bf16-rounded inputs and prints an error map on failure.

    python tools/gemm_selftest.py            # all cases, summary to stdout and gpurun_out/gemm_selftest.txt
    python tools/gemm_selftest.py --case NAME
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _err_report(name, got, ref, tol):
    import torch

    got = got.float()
    ref = ref.float()
    diff = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel = diff.max().item() / denom
    ok = bool(torch.isfinite(got).all()) and rel <= tol
    print(f"[{name}] max|diff|={diff.max().item():.4e} max|ref|={denom:.4e} rel={rel:.3e} tol={tol:.1e} "
          f"{'PASS' if ok else 'FAIL'}")
    if not ok:
        bad = diff > tol * denom
        print(f"   bad elements: {int(bad.sum())} / {bad.numel()}  nonfinite: {int((~torch.isfinite(got)).sum())}")
        flat = bad.reshape(-1, bad.shape[-1])
        rows = flat.any(1).nonzero().flatten()
        cols = flat.any(0).nonzero().flatten()
        print(f"   bad rows (first 24 of {rows.numel()}): {rows[:24].tolist()}")
        print(f"   bad cols (first 24 of {cols.numel()}): {cols[:24].tolist()}")
        g2 = got.reshape(-1, got.shape[-1])
        r2 = ref.reshape(-1, ref.shape[-1])
        print("   got[0,:8]", g2[0, :8].tolist())
        print("   ref[0,:8]", r2[0, :8].tolist())
        if rows.numel():
            i = int(rows[0])
            print(f"   got[{i},:8]", g2[i, :8].tolist())
            print(f"   ref[{i},:8]", r2[i, :8].tolist())
    return ok


def run_case(name: str) -> bool:
    import torch
    import torch.nn.functional as F

    from toist_b200 import kernels as K
    from toist_b200._lib import ACT_GELU, ACT_NONE, ACT_RELU, GEMM_DGRAD, GEMM_FWD, GEMM_WGRAD

    torch.manual_seed(0)
    dev = "cuda"
    bf = torch.bfloat16

    def rn(*s, scale=1.0):
        return (torch.randn(*s, device=dev) * scale).to(bf)

    if name.startswith("linear_fwd"):
        cfgs = {"linear_fwd_basic": (256, 256, 128), "linear_fwd_ragged": (300, 200, 264),
                "linear_fwd_small_n": (800, 256, 4), "linear_fwd_k32": (416, 32, 416),
                "linear_fwd_big": (3328, 2048, 256), "linear_fwd_wide": (3328, 256, 2048)}
        M, Kd, N = cfgs[name]
        x, w = rn(M, Kd), rn(N, Kd, scale=Kd ** -0.5)
        bias = torch.randn(N, device=dev)
        res = rn(M, N)
        ok = True
        y = K.linear_fwd(x, w, bias, act=ACT_RELU, res=res)
        ref = F.relu(x.float() @ w.float().t() + bias + res.float())
        ok &= _err_report(name + "/bf16+bias+res+relu", y, ref, 1e-2)
        y32 = K.linear_fwd(x, w, None, out_dtype=torch.float32)
        ok &= _err_report(name + "/f32 plain", y32, x.float() @ w.float().t(), 2e-5)
        aux = torch.empty(M, N, device=dev, dtype=bf)
        yg = K.linear_fwd(x, w, bias, act=ACT_GELU, out_dtype=torch.float32, aux=aux)
        pre = x.float() @ w.float().t() + bias
        ok &= _err_report(name + "/gelu f32", yg, F.gelu(pre), 2e-5)
        ok &= _err_report(name + "/aux preact", aux, pre, 1e-2)
        return ok
    if name.startswith("linear_dgrad"):
        M, N, Kd = {"linear_dgrad_basic": (256, 128, 256), "linear_dgrad_ragged": (300, 264, 200),
                    "linear_dgrad_big": (3328, 2048, 256)}[name]
        dy, w = rn(M, N), rn(N, Kd, scale=N ** -0.5)
        act = rn(M, Kd)
        dx = K.linear_dgrad(dy, w, mask=act, out_dtype=torch.float32)
        ref = (dy.float() @ w.float()) * (act.float() > 0)
        return _err_report(name, dx, ref, 2e-5)
    if name.startswith("linear_wgrad"):
        M, N, Kd = {"linear_wgrad_basic": (256, 128, 256), "linear_wgrad_ragged": (300, 264, 200),
                    "linear_wgrad_big": (3328, 256, 2048)}[name]
        dy, x = rn(M, N, scale=M ** -0.5), rn(M, Kd)
        dw = torch.zeros(N, Kd, device=dev)
        K.linear_wgrad(dy, x, dw, accumulate=True)
        ok = _err_report(name + "/atomic", dw, dy.float().t() @ x.float(), 2e-5)
        dw2 = torch.full((N, Kd), 7.0, device=dev)
        K.linear_wgrad(dy, x, dw2, accumulate=False)
        ok &= _err_report(name + "/store", dw2, dy.float().t() @ x.float(), 2e-5)
        return ok
    if name.startswith("conv_"):
        # name: conv_<fwd|dgrad|wgrad>_<k>x<k>_s<stride>_<H>  e.g. conv_fwd_3_s1_40
        _, kind, ks, st, hs = name.split("_")
        k, stride, H = int(ks), int(st[1:]), int(hs)
        pad = k // 2
        n, cin, cout = 2, 64, 128
        if H <= 24:
            n, cin, cout = 8, 128, 192
        x = rn(n, H, H, cin)
        w = rn(cout, k, k, cin, scale=(k * k * cin) ** -0.5)
        x_nchw = x.float().permute(0, 3, 1, 2)
        w_oihw = w.float().permute(0, 3, 1, 2)
        if kind == "fwd":
            shift = torch.randn(cout, device=dev)
            y = K.conv_fwd(x, w, shift, stride=stride, pad=pad, act=ACT_RELU, out_dtype=torch.float32)
            ref = F.relu(F.conv2d(x_nchw, w_oihw, shift, stride=stride, padding=pad)).permute(0, 2, 3, 1)
            return _err_report(name, y, ref, 3e-5)
        Ho = (H + 2 * pad - k) // stride + 1
        dy = rn(n, Ho, Ho, cout, scale=0.1)
        dy_nchw = dy.float().permute(0, 3, 1, 2)
        if kind == "dgrad":
            dx = K.conv_dgrad(dy, w, (H, H), stride=stride, pad=pad, mask=x)
            ref = torch.nn.grad.conv2d_input((n, cin, H, H), w_oihw, dy_nchw, stride=stride, padding=pad)
            ref = (ref * (x_nchw > 0)).permute(0, 2, 3, 1)
            return _err_report(name, dx, ref, 1e-2)
        if kind == "wgrad":
            dw = torch.zeros(cout, k, k, cin, device=dev)
            K.conv_wgrad(dy, x, dw, stride=stride, pad=pad)
            ref = torch.nn.grad.conv2d_weight(x_nchw, (cout, cin, k, k), dy_nchw, stride=stride, padding=pad)
            return _err_report(name, dw, ref.permute(0, 2, 3, 1), 3e-5)
    if name.startswith("epi_"):
        # bf16 outputs: the shared-memory / TMA epilogue (res and mask tiles by TMA load, TMA store with clipping)
        ok = True
        if name == "epi_linear_ragged":
            for (M, Kd, N) in ((300, 200, 264), (128, 64, 8), (1000, 512, 1032), (77, 256, 256)):
                x, w = rn(M, Kd), rn(N, Kd, scale=Kd ** -0.5)
                bias, res = torch.randn(N, device=dev), rn(M, N)
                y = K.linear_fwd(x, w, bias, act=ACT_RELU, res=res)
                ref = F.relu(x.float() @ w.float().t() + bias + res.float()).to(bf).float()
                ok &= _err_report(f"{name}/{M}x{Kd}x{N} fwd+res", y, ref, 8e-3)
                dy, act = rn(M, N), rn(M, Kd)
                w2 = rn(N, Kd, scale=N ** -0.5)
                dx = K.linear_dgrad(dy, w2, mask=act, res=act)
                ref = ((dy.float() @ w2.float() + act.float()) * (act.float() > 0)).to(bf).float()
                ok &= _err_report(f"{name}/{M}x{N}x{Kd} dgrad+res+mask", dx, ref, 8e-3)
            return ok
        if name == "epi_conv":
            for (n, H, cin, cout, k, stride) in ((8, 40, 256, 1024, 1, 1), (8, 40, 256, 512, 3, 1), (3, 21, 64, 256, 1, 1),
                                                 (2, 40, 128, 256, 3, 2), (8, 20, 512, 2048, 1, 1)):
                pad = k // 2
                x = rn(n, H, H, cin)
                w = rn(cout, k, k, cin, scale=(k * k * cin) ** -0.5)
                Ho = (H + 2 * pad - k) // stride + 1
                shift, res = torch.randn(cout, device=dev), rn(n, Ho, Ho, cout)
                y = K.conv_fwd(x, w, shift, stride=stride, pad=pad, act=ACT_RELU, res=res)
                x_nchw, w_oihw = x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2)
                ref = F.relu(F.conv2d(x_nchw, w_oihw, shift, stride=stride, padding=pad).permute(0, 2, 3, 1) + res.float())
                ok &= _err_report(f"{name}/fwd {n}x{H}x{cin}->{cout} k{k}s{stride}", y, ref.to(bf).float(), 8e-3)
                dy = rn(n, Ho, Ho, cout, scale=0.1)
                gi = rn(n, H, H, cin, scale=0.1)
                dx = K.conv_dgrad(dy, w, (H, H), stride=stride, pad=pad, mask=x, res=gi)
                ref = torch.nn.grad.conv2d_input((n, cin, H, H), w_oihw, dy.float().permute(0, 3, 1, 2), stride=stride,
                                                 padding=pad).permute(0, 2, 3, 1)
                ref = ((ref + gi.float()) * (x.float() > 0)).to(bf).float()
                ok &= _err_report(f"{name}/dgrad {n}x{H}x{cout}->{cin} k{k}s{stride}", dx, ref, 8e-3)
            return ok
    if name == "attn_batched":
        # scores[b,h,q,k] = Q[q,b,h,:] . K[k,b,h,:]   from a packed [S, B, 3, H, D] projection (D = 32)
        S, B, Hh, D = 416, 2, 8, 32
        qkv = rn(S, B, 3, Hh, D)
        out = torch.empty(B, Hh, S, S, device=dev, dtype=torch.float32)
        sS, sB, s3, sH = qkv.stride(0), qkv.stride(1), qkv.stride(2), qkv.stride(3)
        a = K.t4(qkv, (D, S, Hh, B), (1, sS, sH, sB), offset=0)
        b = K.t4(qkv, (D, S, Hh, B), (1, sS, sH, sB), offset=s3)
        K.gemm(GEMM_FWD, a, b, out, ext=(S, Hh, B), tile=(128, 1, 1), n_cols=S, out_strides=(S, S * S, Hh * S * S),
               k_per_tap=D, b_batched=True, alpha=D ** -0.5)
        q = qkv[:, :, 0].float().permute(1, 2, 0, 3)
        kk = qkv[:, :, 1].float().permute(1, 2, 0, 3)
        ref = (q @ kk.transpose(-1, -2)) * D ** -0.5
        ok = _err_report(name + "/QK^T", out, ref, 2e-5)
        # O[q,b,h,:] = P[b,h,q,:] @ V[:,b,h,:]  (DGRAD mode: B = V is MN-major, reduction over keys)
        P = torch.softmax(ref, -1).to(bf)
        o = torch.empty(S, B, Hh, D, device=dev, dtype=torch.float32)
        a = K.t4(P, (S, S, Hh, B), (1, S, S * S, Hh * S * S))
        b = K.t4(qkv, (D, S, Hh, B), (1, sS, sH, sB), offset=2 * s3)
        K.gemm(GEMM_DGRAD, a, b, o, ext=(S, Hh, B), tile=(128, 1, 1), n_cols=D,
               out_strides=(o.stride(0), o.stride(2), o.stride(1)), k_per_tap=S, b_batched=True)
        v = qkv[:, :, 2].float().permute(1, 2, 0, 3)
        ref_o = (P.float() @ v).permute(2, 0, 1, 3)
        ok &= _err_report(name + "/PV", o, ref_o, 2e-5)
        # dV[k,b,h,:] = P[b,h,:,k]^T @ dO[:,b,h,:]  (WGRAD mode, batched over (h, b))
        dO = rn(S, B, Hh, D)
        dv = torch.zeros(S, B, Hh, D, device=dev, dtype=torch.float32)
        a = K.t4(P, (S, S, Hh, B), (1, S, S * S, Hh * S * S))
        b = K.t4(dO, (D, S, Hh, B), (1, dO.stride(0), dO.stride(2), dO.stride(1)))
        K.gemm(GEMM_WGRAD, a, b, dv, ext=(S, 1, 1), tile=(64, 1, 1), n_cols=D, m_rows=S,
               out_strides=(dv.stride(0), dv.stride(2), dv.stride(1)), batch=(Hh, B))
        ref_dv = (P.float().transpose(-1, -2) @ dO.float().permute(1, 2, 0, 3)).permute(2, 0, 1, 3)
        ok &= _err_report(name + "/dV", dv, ref_dv, 2e-5)
        return ok
    raise SystemExit(f"unknown case {name}")


CASES = [
    "linear_fwd_basic", "linear_fwd_ragged", "linear_fwd_small_n", "linear_fwd_k32", "linear_fwd_big",
    "linear_fwd_wide",
    "linear_dgrad_basic", "linear_dgrad_ragged", "linear_dgrad_big",
    "linear_wgrad_basic", "linear_wgrad_ragged", "linear_wgrad_big",
    "conv_fwd_1_s1_40", "conv_fwd_3_s1_40", "conv_fwd_3_s1_20", "conv_fwd_3_s2_40", "conv_fwd_1_s2_40",
    "conv_fwd_7_s2_40",
    "conv_dgrad_3_s1_40", "conv_dgrad_3_s1_20", "conv_dgrad_3_s2_40", "conv_dgrad_1_s2_40",
    "conv_wgrad_3_s1_40", "conv_wgrad_3_s1_20", "conv_wgrad_3_s2_40", "conv_wgrad_1_s1_40",
    "attn_batched", "epi_linear_ragged", "epi_conv",
]


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", default="")
    ap.add_argument("--inproc", action="store_true", help="run every case in this process (fast; a trap kills all)")
    args = ap.parse_args()
    if args.case:
        return 0 if run_case(args.case) else 1
    if args.inproc:
        bad = [c for c in CASES if (not args.only or args.only in c) and not run_case(c)]
        print(f"gemm_selftest: {len(bad)} failed {bad}")
        return 1 if bad else 0
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    lines = []
    fails = 0
    for c in CASES:
        if args.only and args.only not in c:
            continue
        try:
            r = subprocess.run([sys.executable, __file__, "--case", c], capture_output=True, text=True, timeout=300)
            txt = r.stdout + ("\n" + r.stderr[-3000:] if r.returncode != 0 else "")
            status = "PASS" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
        except subprocess.TimeoutExpired:
            txt, status = "", "TIMEOUT"
        if status != "PASS":
            fails += 1
        lines.append(f"=== {c}: {status}\n{txt}")
        print(lines[-1], flush=True)
        (out_dir / "gemm_selftest.txt").write_text("\n".join(lines))
    print(f"gemm_selftest: {len(CASES) - fails} passed, {fails} failed")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
