"""Tensor-level wrappers over the C ABI (no autograd here; see functional.py for the autograd.Functions).

Every function launches hand-written sm_100a kernels from libtoist_b200.so on torch's current CUDA stream.
torch is used only to own device memory.  Layout conventions:
  * activations of the convolutional trunk are NHWC ("channels last") bf16,
  * sequence activations are [rows, features] bf16 with features contiguous,
  * weights are bf16 [out_features, in_features] (conv: [Cout, kh, kw, Cin]), master copies stay fp32 nn.Parameters.
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, BF16, F32, GEMM_DGRAD, GEMM_FWD, GEMM_WGRAD, GemmDesc,
                   Tap, Tensor4)

_launch_count = 0
_gemm_profiler = None  # bench.py installs an object with .record(tag, flops) -> context manager around the launch
_gemm_tag = "gemm"


def set_gemm_profiler(p) -> None:
    global _gemm_profiler
    _gemm_profiler = p


class _NoProf:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def _prof(tag: str, flops: float = 0.0, sig=None, relaunch=None):
    """`sig` identifies the launch shape, `relaunch()` issues the same launch again on the current stream: bench.py
    replays every distinct launch in isolation to get per-kernel durations free of host launch gaps."""
    return _NoProf() if _gemm_profiler is None else _gemm_profiler.record(tag, flops, sig, relaunch)


class gemm_tag:
    """`with gemm_tag("attn_core"):` labels the GEMM launches issued inside (used for the per-family roofline)."""

    def __init__(self, tag: str):
        self.tag = tag

    def __enter__(self):
        global _gemm_tag
        self.prev, _gemm_tag = _gemm_tag, self.tag

    def __exit__(self, *exc):
        global _gemm_tag
        _gemm_tag = self.prev


def launches() -> int:
    """Number of toist_b200 kernel launches issued by this process (bench.py reports the per-step delta)."""
    return _launch_count


def _count(n: int = 1) -> None:
    global _launch_count
    _launch_count += n


class _WgradLane:
    """Second stream for parameter-gradient kernels.  Inside a backward stage (StageFn.backward) the weight-gradient
    GEMMs, bias column sums and their layout permutes do not feed the data-gradient chain, so they are issued on a side
    stream: the GPU then always has a second kernel's CTAs to fill the SMs the (short, 100..300-CTA) data-gradient
    kernels leave idle, and their launch / prologue / epilogue latencies overlap.  Works in eager mode and inside
    CUDA-graph capture (fork = event wait, join = `join()` before the stage returns).  TOIST_WGRAD_LANE=0 disables."""

    enabled = os.environ.get("TOIST_WGRAD_LANE", "1") != "0"
    active = False   # only StageFn.backward turns the lane on; direct block calls stay single-stream
    stream = None
    dirty = False
    keep: list = []
    also_wait: list = []  # further streams producing the lane's inputs (the trunk's second data-gradient chain)


class wgrad_lanes:
    """`with wgrad_lanes():` around a backward stage body; joins the side stream on exit."""

    def __enter__(self):
        self.prev, _WgradLane.active = _WgradLane.active, _WgradLane.enabled
        return self

    def __exit__(self, *exc):
        _WgradLane.active = self.prev
        wgrad_join()
        return False


class wgrad_lane:
    """`with wgrad_lane(dy, x):` issues the enclosed launches on the side stream (after everything already queued on
    the current stream) and keeps the named tensors alive until the join."""

    def __init__(self, *tensors):
        self.tensors = tensors
        self.ctx = None

    def __enter__(self):
        L = _WgradLane
        if not L.active:
            return self
        main = torch.cuda.current_stream()
        if L.stream is None or L.stream.device != main.device:
            L.stream = torch.cuda.Stream(device=main.device)
        L.stream.wait_stream(main)
        for s_ in L.also_wait:
            L.stream.wait_stream(s_)
        L.keep.extend(t for t in self.tensors if t is not None)
        L.dirty = True
        self.ctx = torch.cuda.stream(L.stream)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
            self.ctx = None
        return False


class gemm_chains:
    """`with gemm_chains(n):` tells the engine that n independent launch chains run side by side (tile-shape hint)."""

    def __init__(self, n: int):
        self.n = n

    def __enter__(self):
        self.prev = _lib.load().toist_gemm_concurrency(self.n)
        return self

    def __exit__(self, *exc):
        _lib.load().toist_gemm_concurrency(self.prev)
        return False


def keep_alive(*tensors) -> None:
    """Holds the tensors until the next `wgrad_join()` (end of the backward stage): buffers that a side stream still
    reads must not go back to the allocator of the stream that made them."""
    _WgradLane.keep.extend(t for t in tensors if t is not None)


def wgrad_join() -> None:
    L = _WgradLane
    if L.dirty:
        torch.cuda.current_stream().wait_stream(L.stream)
        L.dirty = False
    L.keep.clear()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def t4(t: torch.Tensor, dims: Sequence[int], strides: Sequence[int], offset: int = 0) -> Tensor4:
    """A 4-D bf16 view (dim[0] innermost) over `t`'s storage, starting `offset` elements after t.data_ptr()."""
    assert t.dtype == torch.bfloat16, "GEMM operands must be bf16"
    r = Tensor4()
    r._keep = t  # the profiler's relaunch closures must keep the operand storage alive (see gemm())
    r.ptr = t.data_ptr() + 2 * offset
    for i in range(4):
        r.dim[i] = int(dims[i])
        r.stride[i] = int(strides[i])
    return r


@lru_cache(maxsize=4096)
def pick_tile(ext_x: int, ext_y: int, ext_n: int, rows: int = 128, max_x: int = 128) -> Tuple[int, int, int]:
    """Chooses a (tile_x, tile_y, tile_n) box holding `rows` pixels that wastes the fewest padded rows."""
    best = None
    tx = 1
    while tx <= min(rows, max_x):
        ty = 1
        while tx * ty <= rows:
            tn = rows // (tx * ty)
            if tx * ty * tn == rows and tn <= 256 and ty <= 256:
                padded = (-(-ext_x // tx)) * (-(-ext_y // ty)) * (-(-ext_n // tn)) * rows
                key = (padded, -tx, -ty)
                if best is None or key < best[0]:
                    best = (key, (tx, ty, tn))
            ty *= 2
        tx *= 2
    return best[1]


_gemm_ws = {}  # device index -> zeroed scheduler workspace of the persistent GEMM kernel (kept alive for the process)


def _ensure_gemm_workspace(dev: torch.device) -> None:
    """Hands libtoist_b200 the tile-scheduler counters of its persistent GEMM kernel once per device: the library never
    allocates device memory itself (toist_gemm_set_workspace).  512 KB = 65536 launch slots."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx in _gemm_ws:
        return
    if torch.cuda.is_current_stream_capturing():  # allocate outside of captures; until then the one-tile kernel runs
        return
    with torch.cuda.device(idx):
        ws = torch.zeros(128 * 1024, dtype=torch.int32, device=torch.device("cuda", idx))
        torch.cuda.synchronize(idx)
        _lib.check(_lib.load().toist_gemm_set_workspace(ws.data_ptr(), ws.numel() * 4))
    _gemm_ws[idx] = ws


def gemm(mode: int, a: Tensor4, b: Tensor4, out: torch.Tensor, *, ext: Tuple[int, int, int],
         tile: Tuple[int, int, int], n_cols: int, out_strides: Tuple[int, int, int], k_per_tap: int = 0,
         taps: Sequence[Tuple[int, int, int, int]] = ((0, 0, 0, 0),), stride: Tuple[int, int] = (1, 1),
         m_rows: int = 0, b_batched: bool = False, batch: Tuple[int, int] = (1, 1), splits: int = 1,
         out_offset: int = 0, alpha: float = 1.0, col_scale: Optional[torch.Tensor] = None,
         col_shift: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
         res: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
         act: int = ACT_NONE, accumulate: bool = False) -> None:
    """Raw launch of the implicit-GEMM engine; see include/toist_b200.h for the contract.

    `res`, `mask`, `aux` share the output's addressing (out_strides / out_offset).
    """
    _ensure_gemm_workspace(out.device)
    d = GemmDesc()
    d.mode = mode
    d.a = a
    d.b = b
    d.ext_x, d.ext_y, d.ext_n = ext
    d.tile_x, d.tile_y, d.tile_n = tile
    d.stride_x, d.stride_y = stride
    d.n_cols = n_cols
    d.m_rows = m_rows
    d.k_per_tap = k_per_tap
    d.n_taps = len(taps)
    for i, (dx, dy, dn, col) in enumerate(taps):
        d.taps[i].dx, d.taps[i].dy, d.taps[i].dn, d.taps[i].col = dx, dy, dn, col
    d.b_batched = 1 if b_batched else 0
    d.batch_y, d.batch_n = batch
    d.splits = splits
    esz = out.element_size()
    d.out = out.data_ptr() + esz * out_offset
    d.out_dtype = _dt(out)
    d.out_sx, d.out_sy, d.out_sn = out_strides
    d.alpha = alpha
    for name, t in (("col_scale", col_scale), ("col_shift", col_shift), ("row_scale", row_scale)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous()
            setattr(d, name, t.data_ptr())
    if res is not None:
        d.res = res.data_ptr() + res.element_size() * out_offset
        d.res_dtype = _dt(res)
    if mask is not None:
        assert mask.dtype == torch.bfloat16
        d.mask = mask.data_ptr() + 2 * out_offset
    if aux is not None:
        assert aux.dtype == torch.bfloat16
        d.aux = aux.data_ptr() + 2 * out_offset
    d.act = act
    d.accumulate = 1 if accumulate else 0
    if _gemm_profiler is None:
        _lib.check(_lib.load().toist_gemm(C.byref(d), _stream()))
    else:
        pix = ext[0] * ext[1] * ext[2]
        if mode == GEMM_WGRAD:
            flops = 2.0 * m_rows * n_cols * len(taps) * pix * batch[0] * batch[1]
        else:
            flops = 2.0 * pix * n_cols * k_per_tap * len(taps)
        sig = ("gemm", mode, tuple(ext), tuple(tile), n_cols, m_rows, k_per_tap, len(taps), tuple(stride), splits,
               tuple(batch), act, d.out_dtype, res is not None, mask is not None, aux is not None, bool(accumulate),
               bool(b_batched), tuple(out_strides))
        # algorithmic bytes of the launch: every operand read once, the output written once (DESIGN.md "Measurement")
        sxy = stride[0] * stride[1] if len(taps) > 1 else 1
        if mode == GEMM_WGRAD:
            abytes = (pix * m_rows * 2 + pix * sxy * n_cols * 2) * batch[0] * batch[1] + m_rows * n_cols * len(taps) * 4
        else:
            abytes = pix * sxy * k_per_tap * 2 + n_cols * k_per_tap * len(taps) * 2 * (pix // 128 if b_batched else 1) \
                + pix * n_cols * (out.element_size() + (res.element_size() if res is not None else 0)
                                  + (2 if mask is not None else 0) + (2 if aux is not None else 0))
        if hasattr(_gemm_profiler, "bytes"):
            _gemm_profiler.bytes[sig] = float(abytes)
        dcopy = GemmDesc.from_buffer_copy(d)
        keep = (out, res, mask, aux, col_scale, col_shift, row_scale, getattr(a, "_keep", None), getattr(b, "_keep", None))

        def relaunch(dcopy=dcopy, keep=keep):
            _lib.check(_lib.load().toist_gemm(C.byref(dcopy), _stream()))

        with _gemm_profiler.record(_gemm_tag, flops, sig, relaunch):
            _lib.check(_lib.load().toist_gemm(C.byref(d), _stream()))
    _count()


# ------------------------------------------------------------------------------------------------ dense layers
def linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE,
               res: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
               aux: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
               alpha: float = 1.0) -> torch.Tensor:
    """y[M,N] = act(alpha * x[M,K] @ w[N,K]^T + bias + res).  x, w bf16 with contiguous last dim."""
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and x.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    assert out.stride(1) == 1
    gemm(GEMM_FWD, t4(x, (K, M, 1, 1), (1, x.stride(0), 0, 0)), t4(w, (K, N, 1, 1), (1, w.stride(0), 0, 0)), out,
         ext=(M, 1, 1), tile=(128, 1, 1), n_cols=N, out_strides=(out.stride(0), 0, 0), k_per_tap=K,
         col_shift=bias, res=res, aux=aux, act=act, alpha=alpha)
    return out


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, *, mask: Optional[torch.Tensor] = None,
                 res: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
                 out: Optional[torch.Tensor] = None, alpha: float = 1.0) -> torch.Tensor:
    """dx[M,K] = (alpha * dy[M,N] @ w[N,K] + res) * (mask > 0)."""
    M, N = dy.shape
    K = w.shape[1]
    assert w.shape[0] == N and dy.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, K), dtype=out_dtype, device=dy.device)
    gemm(GEMM_DGRAD, t4(dy, (N, M, 1, 1), (1, dy.stride(0), 0, 0)), t4(w, (K, N, 1, 1), (1, w.stride(0), 0, 0)), out,
         ext=(M, 1, 1), tile=(128, 1, 1), n_cols=K, out_strides=(out.stride(0), 0, 0), k_per_tap=N, res=res,
         mask=mask, alpha=alpha)
    return out


_WGRAD_CTAS = int(os.environ.get("TOIST_WGRAD_CTAS", "296"))  # split-K target: CTAs per weight-gradient launch


def _wgrad_splits(m_rows: int, n_cols: int, n_taps: int, pixel_tiles: int, batch: int = 1) -> int:
    base = -(-m_rows // 128) * -(-n_cols // 128) * n_taps * batch
    want = max(1, _WGRAD_CTAS // max(base, 1))
    return max(1, min(want, pixel_tiles // 4 if pixel_tiles >= 8 else 1))


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, *, accumulate: bool = True,
                 row_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dw[N,K] (+)= dy[M,N]^T @ x[M,K]  (fp32 output, split over M with atomics when it helps occupancy)."""
    M, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and dw.shape == (N, K) and dw.dtype == torch.float32 and dw.stride(1) == 1
    tiles = -(-M // 64)
    splits = _wgrad_splits(N, K, 1, tiles) if accumulate else 1
    gemm(GEMM_WGRAD, t4(dy, (N, M, 1, 1), (1, dy.stride(0), 0, 0)), t4(x, (K, M, 1, 1), (1, x.stride(0), 0, 0)), dw,
         ext=(M, 1, 1), tile=(64, 1, 1), n_cols=K, m_rows=N, out_strides=(dw.stride(0), 0, 0), splits=splits,
         row_scale=row_scale, accumulate=accumulate)
    return dw


# ------------------------------------------------------------------------------------------------ convolutions
def _nhwc_t4(x: torch.Tensor) -> Tensor4:
    n, h, w, c = x.shape
    assert x.stride(3) == 1
    return t4(x, (c, w, h, n), (1, x.stride(2), x.stride(1), x.stride(0)))


def conv_out_size(size: int, k: int, stride: int, pad: int) -> int:
    return (size + 2 * pad - k) // stride + 1


def conv_fwd(x: torch.Tensor, w: torch.Tensor, shift: Optional[torch.Tensor] = None, *, stride: int = 1, pad: int = 0,
             act: int = ACT_NONE, res: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
             scale: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[N,Ho,Wo,Cout] = act((conv(x[N,H,W,Cin], w[Cout,kh,kw,Cin]) * scale + shift) + res), all NHWC bf16."""
    n, h, wd, cin = x.shape
    cout, kh, kw, cin2 = w.shape
    assert cin == cin2 and w.is_contiguous()
    ho, wo = conv_out_size(h, kh, stride, pad), conv_out_size(wd, kw, stride, pad)
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=out_dtype, device=x.device)
    assert out.shape == (n, ho, wo, cout) and out.is_contiguous() and out.dtype == out_dtype
    taps = [(kx - pad, ky - pad, 0, (ky * kw + kx) * cin) for ky in range(kh) for kx in range(kw)]
    tile = pick_tile(wo, ho, n, 128, 128 // stride if stride > 1 else 128)
    gemm(GEMM_FWD, _nhwc_t4(x), t4(w, (kh * kw * cin, cout, 1, 1), (1, kh * kw * cin, 0, 0)), out,
         ext=(wo, ho, n), tile=tile, n_cols=cout, out_strides=(cout, wo * cout, ho * wo * cout), k_per_tap=cin,
         taps=taps, stride=(stride, stride), col_scale=scale, col_shift=shift, res=res, act=act)
    return out


def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, in_hw: Tuple[int, int], *, stride: int = 1, pad: int = 0,
               mask: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx[N,H,W,Cin] = (conv_transpose(dy[N,Ho,Wo,Cout], w[Cout,kh,kw,Cin]) + res) * (mask > 0)."""
    n, ho, wo, cout = dy.shape  # dy may carry zero-padded channels beyond the filter count (16-byte TMA rows)
    cout2, kh, kw, cin = w.shape
    assert cout >= cout2 and w.is_contiguous()
    h, wd = in_hw
    dx = out if out is not None else torch.empty((n, h, wd, cin), dtype=torch.bfloat16, device=dy.device)
    assert dx.shape == (n, h, wd, cin) and dx.is_contiguous() and dx.dtype == torch.bfloat16
    a = _nhwc_t4(dy)
    # B rows = the filters that exist: the zero-padded channels of dy meet TMA zero fill, never memory past the tensor
    b = t4(w, (kh * kw * cin, cout2, 1, 1), (1, kh * kw * cin, 0, 0))
    for phy in range(stride):
        for phx in range(stride):
            qh, qw = -(-(h - phy) // stride), -(-(wd - phx) // stride)
            if qh <= 0 or qw <= 0:
                continue
            taps = []
            for ky in range(kh):
                if (phy + pad - ky) % stride:
                    continue
                for kx in range(kw):
                    if (phx + pad - kx) % stride:
                        continue
                    taps.append(((phx + pad - kx) // stride, (phy + pad - ky) // stride, 0, (ky * kw + kx) * cin))
            gemm(GEMM_DGRAD, a, b, dx, ext=(qw, qh, n), tile=pick_tile(qw, qh, n), n_cols=cin,
                 out_strides=(stride * cin, stride * wd * cin, h * wd * cin), out_offset=(phy * wd + phx) * cin,
                 k_per_tap=cout, taps=taps, res=res, mask=mask)
    return dx


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, *, stride: int = 1, pad: int = 0,
               accumulate: bool = True, row_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dw[Cout,kh,kw,Cin] (+)= sum_pixels dy[pix, Cout] * x[pix*stride + tap, Cin]  (fp32)."""
    n, ho, wo, cout_dy = dy.shape  # may be zero padded beyond dw's filter count
    n2, h, wd, cin = x.shape
    cout, kh, kw, cin2 = dw.shape
    assert n == n2 and cout_dy >= cout and cin == cin2 and dw.dtype == torch.float32 and dw.is_contiguous()
    taps = [(kx - pad, ky - pad, 0, (ky * kw + kx) * cin) for ky in range(kh) for kx in range(kw)]
    tile = pick_tile(wo, ho, n, 64, 64)
    ptiles = -(-wo // tile[0]) * -(-ho // tile[1]) * -(-n // tile[2])
    splits = _wgrad_splits(cout, cin, len(taps), ptiles) if accumulate else 1
    gemm(GEMM_WGRAD, _nhwc_t4(dy), _nhwc_t4(x), dw, ext=(wo, ho, n), tile=tile, n_cols=cin, m_rows=cout,
         out_strides=(kh * kw * cin, 0, 0), taps=taps, stride=(stride, stride), splits=splits, row_scale=row_scale,
         accumulate=accumulate)
    return dw


# ------------------------------------------------------------------------------------------------ helpers
def _L():
    return _lib.load()


def _ck(rc: int, n: int = 1) -> None:
    _lib.check(rc)
    _count(n)


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 -> bf16 copy (contiguous)."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = out if out is not None else torch.empty_like(x, dtype=torch.bfloat16)
    assert y.dtype == torch.bfloat16 and y.is_contiguous() and y.numel() == x.numel()
    _ck(_L().toist_cast_f32_bf16(x.data_ptr(), y.data_ptr(), x.numel(), _stream()))
    return y


def cast_f32(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    y = out if out is not None else torch.empty_like(x, dtype=torch.float32)
    assert y.dtype == torch.float32 and y.is_contiguous() and y.numel() == x.numel()
    _ck(_L().toist_cast_bf16_f32(x.data_ptr(), y.data_ptr(), x.numel(), _stream()))
    return y


def dropout(x: torch.Tensor, drop, res: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Dropout with a regenerable mask: drop = (p, seed uint64-as-int64 device tensor [1], site id).
    out = keep ? x / (1 - p) : 0 (+ res).  Calling it again on the upstream gradient is the backward pass."""
    p, seed, site = drop
    assert x.is_contiguous() and (res is None or (res.is_contiguous() and res.dtype == x.dtype))
    y = out if out is not None else torch.empty_like(x)
    _ck(_L().toist_dropout(x.data_ptr(), _ptr(res), y.data_ptr(), x.numel(), _dt(x), float(p), seed.data_ptr(),
                           int(site), _stream()))
    return y


def cast_pad_bf16(x: torch.Tensor, ld: int) -> torch.Tensor:
    """fp32 [..., n] -> bf16 [..., ld] zero padded (ld >= n)."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    n = x.shape[-1]
    y = torch.empty((*x.shape[:-1], ld), dtype=torch.bfloat16, device=x.device)
    _ck(_L().toist_cast_pad_f32_bf16(x.data_ptr(), y.data_ptr(), x.numel() // n, n, ld, _stream()))
    return y


def add_bf16(a: torch.Tensor, b: torch.Tensor, c: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert a.dtype == b.dtype == torch.bfloat16 and a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    out = torch.empty_like(a)
    _ck(_L().toist_add_bf16(a.data_ptr(), b.data_ptr(), _ptr(c), out.data_ptr(), a.numel(), _stream()))
    return out


class WeightPrep:
    """One launch that refreshes every bf16 shadow weight from its fp32 master (optionally folding a per-row scale,
    e.g. the FrozenBatchNorm scale of the following norm layer)."""

    def __init__(self, device):
        self.device = device
        self.items = []  # (src, dst, scale, rows, cols, ldd)
        self._table = None
        self._blocks = 0
        self._keep = []

    def add(self, src: torch.Tensor, dst: torch.Tensor, rows: int, cols: int, ldd: Optional[int] = None,
            row_scale: Optional[torch.Tensor] = None, taps: int = 1) -> None:
        """`taps` > 1: src is a conv weight [rows, cols // taps, taps] (OIHW) written as [rows, taps, cols // taps]."""
        assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.is_contiguous()
        assert src.numel() == rows * cols and cols % taps == 0
        self.items.append((src, dst, row_scale, rows, cols, ldd or cols, taps))
        self._table = None

    def _build(self) -> None:
        import struct

        buf = bytearray()
        blocks = 0
        for src, dst, sc, rows, cols, ldd, taps in self.items:
            buf += struct.pack("<QQQiiiiii", src.data_ptr(), dst.data_ptr(), 0 if sc is None else sc.data_ptr(), rows,
                               cols, ldd, blocks, taps, 0)
            blocks += -(-(rows * cols) // 2048)
        host = torch.frombuffer(buf, dtype=torch.uint8).clone()
        self._table = host.to(self.device)
        self._blocks = blocks
        self._ptrs = [(s.data_ptr(), d.data_ptr()) for s, d, *_ in self.items]

    def run(self) -> None:
        if not self.items:
            return
        if self._table is None or self._ptrs != [(s.data_ptr(), d.data_ptr()) for s, d, *_ in self.items]:
            self._build()
        _ck(_L().toist_weight_prep(self._table.data_ptr(), len(self.items), self._blocks, _stream()))


def stem_im2col(images: torch.Tensor, ldk: int = 192) -> torch.Tensor:
    n, c, h, w = images.shape
    assert c == 3 and images.dtype == torch.float32 and images.is_contiguous()
    ho, wo = conv_out_size(h, 7, 2, 3), conv_out_size(w, 7, 2, 3)
    out = torch.empty((n * ho * wo, ldk), dtype=torch.bfloat16, device=images.device)
    _ck(_L().toist_stem_im2col(images.data_ptr(), out.data_ptr(), n, h, w, ldk, _stream()))
    return out


def stem_conv7x7(images: torch.Tensor, w7: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """relu(conv7x7 stride 2 pad 3 (images) * bn_scale + bn_shift) as an implicit GEMM without an im2col matrix.
    images fp32 NCHW [n, 3, h, w]; w7 bf16 [64, 7 * 64]: row o, tap ky, then 8 kernel columns x 8 channels (kx = 7 and
    channels 3..7 zero), BatchNorm scale folded; returns NHWC bf16 [n, ho, wo, 64]."""
    n, c, h, w = images.shape
    assert c == 3 and images.dtype == torch.float32 and images.is_contiguous() and w7.shape == (64, 448)
    ho, wo = conv_out_size(h, 7, 2, 3), conv_out_size(w, 7, 2, 3)
    hp, wp = max(h + 6, 2 * (ho - 1) + 7), max(w + 6, 2 * (wo - 1) + 8)
    xp = torch.empty((n, hp, wp, 8), dtype=torch.bfloat16, device=images.device)
    _ck(_L().toist_stem_pad_nhwc8(images.data_ptr(), xp.data_ptr(), n, h, w, hp, wp, _stream()))
    y = torch.empty((n, ho, wo, 64), dtype=torch.bfloat16, device=images.device)
    # A: "row" x of the map = the 64 elements starting at padded pixel 2 x of an image row (rows overlap by 48 elements)
    a = t4(xp, (64, wo, hp, n), (1, 16, wp * 8, hp * wp * 8))
    b = t4(w7, (448, 64, 1, 1), (1, 448, 0, 0))
    taps = [(0, ky, 0, ky * 64) for ky in range(7)]
    gemm(GEMM_FWD, a, b, y, ext=(wo, ho, n), tile=pick_tile(wo, ho, n, 128, 128), n_cols=64,
         out_strides=(64, wo * 64, ho * wo * 64), k_per_tap=64, taps=taps, stride=(1, 2), col_shift=shift, act=ACT_RELU)
    return y


def maxpool3x3s2(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, 3, 2, 1), conv_out_size(w, 3, 2, 1)
    y = out if out is not None else torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=x.device)
    assert y.shape == (n, ho, wo, c) and y.is_contiguous() and x.is_contiguous()
    _ck(_L().toist_maxpool3x3s2(x.data_ptr(), y.data_ptr(), n, h, w, c, _stream()))
    return y


def colsum(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[c] += sum_r x[r, c] (bias gradient)."""
    assert x.dim() == 2 and x.stride(1) == 1 and out.dtype == torch.float32
    _ck(_L().toist_colsum(x.data_ptr(), _dt(x), out.data_ptr(), x.shape[0], x.shape[1], x.stride(0), _stream()))
    return out


def gelu_bwd(dy: torch.Tensor, pre: torch.Tensor) -> torch.Tensor:
    dx = torch.empty_like(dy)
    _ck(_L().toist_gelu_bwd(dy.data_ptr(), pre.data_ptr(), dx.data_ptr(), dy.numel(), _stream()))
    return dx


def sigmoid_bwd(dy: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    dx = torch.empty_like(dy)
    _ck(_L().toist_sigmoid_bwd(dy.data_ptr(), y.data_ptr(), dx.data_ptr(), dy.numel(), _stream()))
    return dx


def sum_mid(x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    a, r, c = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty((a, c), dtype=torch.float32, device=x.device)
    _ck(_L().toist_sum_mid(x.data_ptr(), _dt(x), out.data_ptr(), a, r, c, 1 if accumulate else 0, _stream()))
    return out


def bcast_mid(x: torch.Tensor, r: int) -> torch.Tensor:
    a, c = x.shape
    out = torch.empty((a, r, c), dtype=torch.bfloat16, device=x.device)
    _ck(_L().toist_bcast_mid(x.data_ptr(), out.data_ptr(), a, r, c, _stream()))
    return out


def nchw_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    n, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    _ck(_L().toist_nchw_to_nhwc(x.data_ptr(), y.data_ptr(), n, c, h * w, _stream()))
    return y


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    _ck(_L().toist_nhwc_to_nchw(x.data_ptr(), y.data_ptr(), n, c, h * w, _stream()))
    return y


# ------------------------------------------------------------------------------------------------ norms / softmax
def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, want_f32: bool = False,
                  want_bf16: bool = True, save_stats: bool = True, out16: Optional[torch.Tensor] = None,
                  out32: Optional[torch.Tensor] = None):
    """x [rows, N] (fp32 or bf16) -> (y_bf16 | None, y_f32 | None, mean, rstd).  out16 / out32: preallocated
    contiguous [rows, N] destinations (e.g. a row slice of a larger buffer)."""
    rows, n = x.shape
    assert x.is_contiguous()
    y16 = out16 if out16 is not None else (
        torch.empty((rows, n), dtype=torch.bfloat16, device=x.device) if want_bf16 else None)
    y32 = out32 if out32 is not None else (
        torch.empty((rows, n), dtype=torch.float32, device=x.device) if want_f32 else None)
    for t in (y16, y32):
        assert t is None or (t.is_contiguous() and t.shape == (rows, n))
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    _ck(_L().toist_layernorm_fwd(x.data_ptr(), _dt(x), gamma.data_ptr(), beta.data_ptr(), _ptr(y16), _ptr(y32),
                                 _ptr(mean), _ptr(rstd), rows, n, eps, _stream()))
    return y16, y32, mean, rstd


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor, *,
                  dy2: Optional[torch.Tensor] = None, dgamma: Optional[torch.Tensor] = None,
                  dbeta: Optional[torch.Tensor] = None, dx_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    rows, n = x.shape
    assert dy.is_contiguous() and x.is_contiguous() and (dy2 is None or (dy2.dtype == dy.dtype and dy2.is_contiguous()))
    dx = torch.empty((rows, n), dtype=dx_dtype, device=x.device)
    _ck(_L().toist_layernorm_bwd(dy.data_ptr(), _ptr(dy2), _dt(dy), x.data_ptr(), _dt(x), mean.data_ptr(),
                                 rstd.data_ptr(), gamma.data_ptr(), dx.data_ptr(), _dt(dx), _ptr(dgamma), _ptr(dbeta),
                                 rows, n, _stream()))
    return dx


_FUSED_LN = os.environ.get("TOIST_FUSED_LN", "1") != "0"


def fused_ln_ok(*tensors) -> bool:
    """bf16 rows of 256 / 512 / 768 contiguous elements, 16-byte aligned: the shapes the fused residual-branch kernels take."""
    if not _FUSED_LN:
        return False
    n = None
    for t in tensors:
        if t is None:
            continue
        if t.dtype != torch.bfloat16 or not t.is_contiguous() or t.data_ptr() % 16:
            return False
        n = t.shape[-1]
    return n in (256, 512, 768)


def layernorm_fused_fwd(x: torch.Tensor, res: Optional[torch.Tensor], add: Optional[torch.Tensor], gamma: torch.Tensor,
                        beta: torch.Tensor, eps: float, drop=None, want_sum: bool = True):
    """s = res + dropout(x); y = LayerNorm(s); y_add = y + add  in one launch (csrc/norm.cu).  drop = (p, seed, site) or
    None.  Returns (y, y_add | None, s | None, mean, rstd); s is written when there is a residual or dropout."""
    rows, n = x.shape
    dev = x.device
    y = torch.empty((rows, n), dtype=torch.bfloat16, device=dev)
    y_add = torch.empty((rows, n), dtype=torch.bfloat16, device=dev) if add is not None else None
    need_s = want_sum and (res is not None or drop is not None)
    s = torch.empty((rows, n), dtype=torch.bfloat16, device=dev) if need_s else None
    mean = torch.empty(rows, dtype=torch.float32, device=dev)
    rstd = torch.empty(rows, dtype=torch.float32, device=dev)
    p, seed, site = drop if drop is not None else (0.0, None, 0)
    _ck(_L().toist_layernorm_fused_fwd(x.data_ptr(), _ptr(res), _ptr(add), gamma.data_ptr(), beta.data_ptr(), _ptr(s),
                                       y.data_ptr(), _ptr(y_add), mean.data_ptr(), rstd.data_ptr(), rows, n, float(eps),
                                       float(p), _ptr(seed), int(site), _stream()))
    return y, y_add, (s if need_s else x), mean, rstd


def layernorm_bwd_drop(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor,
                       drop, *, dy2: Optional[torch.Tensor] = None, dgamma: Optional[torch.Tensor] = None,
                       dbeta: Optional[torch.Tensor] = None):
    """LayerNorm backward + the dropout backward of the branch behind it: returns (dx, dropout_bwd(dx))."""
    rows, n = x.shape
    dx = torch.empty((rows, n), dtype=torch.bfloat16, device=x.device)
    dxd = torch.empty((rows, n), dtype=torch.bfloat16, device=x.device)
    p, seed, site = drop
    _ck(_L().toist_layernorm_bwd_drop(dy.data_ptr(), _ptr(dy2), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                      gamma.data_ptr(), dx.data_ptr(), dxd.data_ptr(), _ptr(dgamma), _ptr(dbeta), rows, n,
                                      float(p), seed.data_ptr(), int(site), _stream()))
    return dx, dxd


def l2norm_fwd(x: torch.Tensor, eps: float = 1e-12):
    rows, n = x.shape
    y = torch.empty_like(x)
    nrm = torch.empty(rows, dtype=torch.float32, device=x.device)
    _ck(_L().toist_l2norm_fwd(x.data_ptr(), y.data_ptr(), nrm.data_ptr(), rows, n, eps, _stream()))
    return y, nrm


def l2norm_bwd(dy: torch.Tensor, y: torch.Tensor, nrm: torch.Tensor) -> torch.Tensor:
    rows, n = y.shape
    dx = torch.empty_like(y)
    _ck(_L().toist_l2norm_bwd(dy.data_ptr(), y.data_ptr(), nrm.data_ptr(), dx.data_ptr(), rows, n, _stream()))
    return dx


def pos_sine(mask_u8: torch.Tensor, num_pos_feats: int, temperature: float = 10000.0, extra_rows: int = 0):
    """mask [B,H,W] uint8 -> (pos_f32, pos_bf16) each [H*W + extra_rows, B, 2*num_pos_feats]; the extra (text) rows
    are zero (models/transformer.py:148)."""
    b, h, w = mask_u8.shape
    assert mask_u8.dtype == torch.uint8 and mask_u8.is_contiguous()
    alloc = torch.zeros if extra_rows else torch.empty
    p32 = alloc((h * w + extra_rows, b, 2 * num_pos_feats), dtype=torch.float32, device=mask_u8.device)
    p16 = alloc((h * w + extra_rows, b, 2 * num_pos_feats), dtype=torch.bfloat16, device=mask_u8.device)
    _ck(_L().toist_pos_sine(mask_u8.data_ptr(), p32.data_ptr(), p16.data_ptr(), b, h, w, num_pos_feats, temperature,
                            _stream()))
    return p32, p16


def embed_gather(ids: torch.Tensor, word: torch.Tensor, pos: torch.Tensor, type_w: torch.Tensor, pad_id: int = 1,
                 seq_first: bool = False):
    """RoBERTa embedding sum; ids [B, L] -> rows b*L + l (or l*B + b when seq_first)."""
    b, l = ids.shape
    e = word.shape[1]
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    out = torch.empty((b * l, e), dtype=torch.float32, device=ids.device)
    pos_ids = torch.empty(b * l, dtype=torch.int32, device=ids.device)
    _ck(_L().toist_embed_gather(ids.data_ptr(), word.data_ptr(), pos.data_ptr(), type_w.data_ptr(), out.data_ptr(),
                                pos_ids.data_ptr(), b, l, e, pad_id, 1 if seq_first else 0, _stream()))
    return out, pos_ids


def embed_scatter(dx: torch.Tensor, ids: torch.Tensor, pos_ids: torch.Tensor, dword: Optional[torch.Tensor],
                  dpos: Optional[torch.Tensor], dtype0: Optional[torch.Tensor], seq_first: bool = False,
                  pad_id: int = -1) -> None:
    b, l = ids.shape
    rows, e = dx.shape
    assert rows == b * l and dx.is_contiguous()
    _ck(_L().toist_embed_scatter(dx.data_ptr(), _dt(dx), ids.data_ptr(), pos_ids.data_ptr(), _ptr(dword), _ptr(dpos),
                                 _ptr(dtype0), b, l, e, 1 if seq_first else 0, int(pad_id), _stream()))


def embed_rows_merge(table: torch.Tensor, ids: torch.Tensor, rows: torch.Tensor, pad_id: int, scale: float) -> None:
    """table[id] = scale * (sum of the rows with that id, in list order), written in place; deterministic."""
    n, e = rows.shape
    assert table.dtype == torch.float32 and table.is_contiguous() and table.shape[1] == e
    assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.numel() == n and rows.is_contiguous()
    _ck(_L().toist_embed_rows_merge(table.data_ptr(), ids.data_ptr(), rows.data_ptr(), n, e, int(pad_id), float(scale),
                                    _stream()))


def permute_021(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [A, B, C] -> [A, C, B] (conv weight gradient OHWI -> OIHW)."""
    a, b, c = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous()
    dst = out if out is not None else torch.empty((a, c, b), dtype=torch.float32, device=src.device)
    assert dst.dtype == torch.float32 and dst.is_contiguous() and dst.numel() == src.numel()
    _ck(_L().toist_permute_021(src.data_ptr(), dst.data_ptr(), a, b, c, _stream()))
    return dst


def key_mask(pad_mask_u8: torch.Tensor, out_hw: Tuple[int, int], text_attention: Optional[torch.Tensor] = None):
    """pad mask [B,H,W] uint8 -> (small [B,h,w] uint8, key [B, h*w + L] uint8); text_attention int64 [B, L]."""
    b, hh, ww = pad_mask_u8.shape
    h, w = out_hw
    assert pad_mask_u8.dtype == torch.uint8 and pad_mask_u8.is_contiguous()
    n_text = 0
    if text_attention is not None:
        assert text_attention.dtype == torch.int64 and text_attention.is_contiguous()
        n_text = text_attention.shape[1]
    small = torch.empty((b, h, w), dtype=torch.uint8, device=pad_mask_u8.device)
    key = torch.empty((b, h * w + n_text), dtype=torch.uint8, device=pad_mask_u8.device)
    _ck(_L().toist_key_mask(pad_mask_u8.data_ptr(), _ptr(text_attention), small.data_ptr(), key.data_ptr(), b, hh, ww,
                            h, w, n_text, _stream()))
    return small, key


# ------------------------------------------------------------------------------------------------ attention (unfused)
def _ld8(n: int) -> int:
    return (n + 7) // 8 * 8


FUSED_ATTENTION = os.environ.get("TOIST_FUSED_ATTN", "1") != "0"


class FusedAttnSaved:
    """What the fused forward keeps for its backward: the output, the row statistics `lse` = (raw row maximum of q.k,
    softmax denominator) and the key mask (the scores and probabilities never leave the SM)."""
    __slots__ = ("ctx", "lse", "key_mask")

    def __init__(self, ctx, lse, key_mask):
        self.ctx, self.lse, self.key_mask = ctx, lse, key_mask


def fused_attention_ok(sq: int, sk: int, d: int) -> bool:
    return FUSED_ATTENTION and _L().toist_attention_supported(sq, sk, d) == 1


def _attn_desc(q, k, v, out, lse, key_mask_u8, nhead, drop) -> "_lib.AttnDesc":
    sq, b, e = q.shape
    a = _lib.AttnDesc()
    for t in (q, k, v, out):
        assert t.dtype == torch.bfloat16 and t.stride(2) == 1
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.lse = lse.data_ptr() if lse is not None else None
    if key_mask_u8 is not None:
        assert key_mask_u8.dtype == torch.uint8 and key_mask_u8.is_contiguous() and key_mask_u8.shape == (b, k.shape[0])
        a.key_mask = key_mask_u8.data_ptr()
    a.q_ss, a.q_sb, a.k_ss, a.k_sb = q.stride(0), q.stride(1), k.stride(0), k.stride(1)
    a.v_ss, a.v_sb, a.o_ss, a.o_sb = v.stride(0), v.stride(1), out.stride(0), out.stride(1)
    a.sq, a.sk, a.b, a.h, a.d = sq, k.shape[0], b, nhead, e // nhead
    if drop is not None:
        p_drop, seed, site = drop
        a.p_drop, a.seed, a.site = float(p_drop), seed.data_ptr(), int(site)
    return a


def attention_fused_fwd(q, k, v, key_mask_u8, nhead: int, ctx: Optional[torch.Tensor] = None, drop=None,
                        need_lse: bool = True):
    """One launch: ctx = dropout(softmax(q k^T / sqrt(d) + mask)) v  (csrc/attention.cu).  Returns (ctx, lse)."""
    sq, b, e = q.shape
    sk = k.shape[0]
    if ctx is None:
        ctx = torch.empty((sq, b, e), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((b, nhead, sq, 2), dtype=torch.float32, device=q.device) if need_lse else None  # (raw row maximum, denominator) per row
    a = _attn_desc(q, k, v, ctx, lse, key_mask_u8, nhead, drop)
    keep = (q, k, v, ctx, lse, key_mask_u8, drop)

    def launch(a=a, keep=keep):
        _lib.check(_L().toist_attention_fwd(C.addressof(a), _stream()))

    with _prof("attn_core", 4.0 * sq * sk * (e // nhead) * nhead * b, ("attn_fwd", sq, sk, b, nhead, e, drop is not None),
               launch):
        launch()
    _count()
    return ctx, lse


def attention_fused_bwd(dctx, q, k, v, saved: FusedAttnSaved, nhead: int, dq, dk, dv, drop=None) -> None:
    sq, b, e = q.shape
    sk = k.shape[0]
    d = e // nhead
    a = _lib.AttnBwdDesc()
    a.fwd = _attn_desc(q, k, v, saved.ctx, saved.lse, saved.key_mask, nhead, drop)
    for t in (dctx, dq, dk, dv):
        assert t.dtype == torch.bfloat16 and t.stride(2) == 1
    ws = torch.empty((_L().toist_attention_bwd_workspace(sq, sk, b, nhead, d) // 4,), dtype=torch.float32, device=q.device)
    a.dout, a.dq, a.dk, a.dv, a.workspace = dctx.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws.data_ptr()
    a.do_ss, a.do_sb, a.dq_ss, a.dq_sb = dctx.stride(0), dctx.stride(1), dq.stride(0), dq.stride(1)
    a.dk_ss, a.dk_sb, a.dv_ss, a.dv_sb = dk.stride(0), dk.stride(1), dv.stride(0), dv.stride(1)
    keep = (dctx, q, k, v, saved, dq, dk, dv, ws, drop)

    def launch(a=a, keep=keep):
        _lib.check(_L().toist_attention_bwd(C.addressof(a), _stream()))

    with _prof("attn_core", 10.0 * sq * sk * d * nhead * b, ("attn_bwd", sq, sk, b, nhead, e, drop is not None), launch):
        launch()
    _count(2)


def attention_dropout_mask(b: int, nhead: int, sq: int, sk: int, drop) -> torch.Tensor:
    """uint8 [b, nhead, sq, sk]: 1 where the fused attention kernels keep the probability (test support)."""
    p_drop, seed, site = drop
    keep = torch.empty((b, nhead, sq, sk), dtype=torch.uint8, device=seed.device)
    _ck(_L().toist_attention_dropout_mask(keep.data_ptr(), b, nhead, sq, sk, float(p_drop), seed.data_ptr(), int(site),
                                          _stream()))
    return keep


def attention_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_mask_u8: Optional[torch.Tensor], nhead: int,
                  need_probs: bool = True, ctx: Optional[torch.Tensor] = None, drop=None, fused: Optional[bool] = None):
    """Multi-head attention core on packed projections.

    q [Sq, B, E], k [Sk, B, E], v [Sk, B, E]: bf16 views whose last dim is contiguous (they may be column slices of a
    wider projection output).  Returns (ctx bf16 [Sq, B, E], probs) where probs is the bf16 [B, H, Sq, ld] softmax
    (None if not needed) or, with dropout (`drop` = (p, seed, site)), the pair (softmax, dropout(softmax)).
    Three launches: QK^T (fp32 scores, scaled), masked softmax (+ dropout), PV.
    """
    sq, b, e = q.shape
    sk = k.shape[0]
    d = e // nhead
    if fused is None:
        fused = fused_attention_ok(sq, sk, d)
    if fused:  # one launch, scores and probabilities stay on the SM; `probs` is then the saved (ctx, lse, mask) record
        ctx, lse = attention_fused_fwd(q, k, v, key_mask_u8, nhead, ctx=ctx, drop=drop, need_lse=need_probs)
        return ctx, (FusedAttnSaved(ctx, lse, key_mask_u8) if need_probs else None)
    ld = _ld8(sk)
    dev = q.device
    scores = torch.empty((b, nhead, sq, ld), dtype=torch.float32, device=dev)
    with gemm_tag("attn_core"):
        return _attention_fwd(q, k, v, key_mask_u8, nhead, need_probs, ctx, scores, sq, sk, b, e, d, ld, dev, drop)


def _attention_fwd(q, k, v, key_mask_u8, nhead, need_probs, ctx, scores, sq, sk, b, e, d, ld, dev, drop):
    gemm(GEMM_FWD, t4(q, (d, sq, nhead, b), (1, q.stride(0), d, q.stride(1))),
         t4(k, (d, sk, nhead, b), (1, k.stride(0), d, k.stride(1))), scores, ext=(sq, nhead, b), tile=(128, 1, 1),
         n_cols=sk, out_strides=(ld, sq * ld, nhead * sq * ld), k_per_tap=d, b_batched=True, alpha=float(d) ** -0.5)
    probs = torch.empty((b, nhead, sq, ld), dtype=torch.bfloat16, device=dev)
    probs_d = torch.empty_like(probs) if drop is not None else None
    p_drop, seed, site = drop if drop is not None else (0.0, None, 0)
    def sm_fwd():
        _ck(_L().toist_attn_softmax_fwd(scores.data_ptr(), _ptr(key_mask_u8), probs.data_ptr(), _ptr(probs_d),
                                        b * nhead * sq, sk, ld, ld, nhead * sq, float(p_drop), _ptr(seed), int(site),
                                        _stream()))

    with _prof("attn_core", 0.0, ("softmax_fwd", b * nhead * sq, sk, probs_d is not None), sm_fwd):
        sm_fwd()
    pv = probs_d if probs_d is not None else probs
    if ctx is None:
        ctx = torch.empty((sq, b, e), dtype=torch.bfloat16, device=dev)
    assert ctx.shape == (sq, b, e) and ctx.stride(2) == 1
    gemm(GEMM_DGRAD, t4(pv, (sk, sq, nhead, b), (1, ld, sq * ld, nhead * sq * ld)),
         t4(v, (d, sk, nhead, b), (1, v.stride(0), d, v.stride(1))), ctx, ext=(sq, nhead, b), tile=(128, 1, 1),
         n_cols=d, out_strides=(ctx.stride(0), d, ctx.stride(1)), k_per_tap=sk, b_batched=True)
    if not need_probs:
        return ctx, None
    return ctx, (probs if probs_d is None else (probs, probs_d))


def attention_bwd(dctx: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, probs, nhead: int,
                  dq: torch.Tensor, dk: torch.Tensor, dv: torch.Tensor, drop=None) -> None:
    """Backward of attention_fwd (`probs` exactly as it returned them, `drop` the same triple).  dq/dk/dv are bf16
    outputs with the same [S, B, E] indexing as q/k/v (they may be column slices of one packed gradient buffer)."""
    if isinstance(probs, FusedAttnSaved):
        attention_fused_bwd(dctx, q, k, v, probs, nhead, dq, dk, dv, drop=drop)
        return
    with gemm_tag("attn_core"):
        _attention_bwd(dctx, q, k, v, probs, nhead, dq, dk, dv, drop)


def _attention_bwd(dctx, q, k, v, probs, nhead, dq, dk, dv, drop) -> None:
    probs_d = None
    if isinstance(probs, (tuple, list)):
        probs, probs_d = probs
    pv = probs_d if probs_d is not None else probs
    p_drop, seed, site = drop if drop is not None else (0.0, None, 0)
    sq, b, e = q.shape
    sk = k.shape[0]
    d = e // nhead
    ld = probs.shape[-1]
    dev = q.device
    sP = (1, ld, sq * ld, nhead * sq * ld)
    # dV[k] = P^T dO      (WGRAD mode: reduction over queries, batched over (h, b))
    gemm(GEMM_WGRAD, t4(pv, (sk, sq, nhead, b), sP),
         t4(dctx, (d, sq, nhead, b), (1, dctx.stride(0), d, dctx.stride(1))), dv, ext=(sq, 1, 1), tile=(64, 1, 1),
         n_cols=d, m_rows=sk, out_strides=(dv.stride(0), d, dv.stride(1)), batch=(nhead, b))
    # dP = dO V^T
    dp = torch.empty((b, nhead, sq, ld), dtype=torch.float32, device=dev)
    gemm(GEMM_FWD, t4(dctx, (d, sq, nhead, b), (1, dctx.stride(0), d, dctx.stride(1))),
         t4(v, (d, sk, nhead, b), (1, v.stride(0), d, v.stride(1))), dp, ext=(sq, nhead, b), tile=(128, 1, 1),
         n_cols=sk, out_strides=(ld, sq * ld, nhead * sq * ld), k_per_tap=d, b_batched=True)
    ds = torch.empty((b, nhead, sq, ld), dtype=torch.bfloat16, device=dev)
    def sm_bwd():
        _ck(_L().toist_attn_softmax_bwd(dp.data_ptr(), probs.data_ptr(), ds.data_ptr(), b * nhead * sq, sk, ld, ld,
                                        float(d) ** -0.5, float(p_drop), _ptr(seed), int(site), _stream()))

    with _prof("attn_core", 0.0, ("softmax_bwd", b * nhead * sq, sk, seed is not None), sm_bwd):
        sm_bwd()
    # dQ = dS K   (DGRAD mode: B = K is MN-major, reduction over keys)
    gemm(GEMM_DGRAD, t4(ds, (sk, sq, nhead, b), sP), t4(k, (d, sk, nhead, b), (1, k.stride(0), d, k.stride(1))), dq,
         ext=(sq, nhead, b), tile=(128, 1, 1), n_cols=d, out_strides=(dq.stride(0), d, dq.stride(1)), k_per_tap=sk,
         b_batched=True)
    # dK = dS^T Q
    gemm(GEMM_WGRAD, t4(ds, (sk, sq, nhead, b), sP), t4(q, (d, sq, nhead, b), (1, q.stride(0), d, q.stride(1))), dk,
         ext=(sq, 1, 1), tile=(64, 1, 1), n_cols=d, m_rows=sk, out_strides=(dk.stride(0), d, dk.stride(1)),
         batch=(nhead, b))


# ------------------------------------------------------------------------------------------------ matcher / criterion
def match_cost(logits: torch.Tensor, boxes: torch.Tensor, tgt_boxes: torch.Tensor, tgt_count: torch.Tensor,
               posmap: torch.Tensor, w_class: float, w_bbox: float, w_giou: float) -> torch.Tensor:
    """logits [L,B,Q,C], boxes [L,B,Q,4] fp32; padded targets -> cost [L,B,Q,Tmax] fp32."""
    L, B, Q, Cc = logits.shape
    tmax = tgt_boxes.shape[1]
    cost = torch.empty((L, B, Q, tmax), dtype=torch.float32, device=logits.device)
    _ck(_L().toist_match_cost(logits.data_ptr(), boxes.data_ptr(), tgt_boxes.data_ptr(), tgt_count.data_ptr(),
                              posmap.data_ptr(), cost.data_ptr(), L, B, Q, Cc, tmax, w_class, w_bbox, w_giou,
                              _stream()))
    return cost


def lsap_device(cost: torch.Tensor, tgt_count: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    L, B, Q, tmax = cost.shape
    match_q = torch.empty((L, B, tmax), dtype=torch.int32, device=cost.device)
    _ck(_L().toist_lsap_device(cost.data_ptr(), tgt_count.data_ptr(), match_q.data_ptr(), flags.data_ptr(), L * B, B, Q,
                               tmax, _stream()))
    return match_q


def lsap_host(cost) -> Tuple["np.ndarray", "np.ndarray"]:
    """scipy.optimize.linear_sum_assignment work-alike on the host (C++ in libtoist_b200.so)."""
    import numpy as np

    c = np.ascontiguousarray(np.asarray(cost, dtype=np.float64))
    if c.ndim != 2:
        raise ValueError("expected a matrix")
    nr, nc = c.shape
    n = min(nr, nc)
    ri = np.zeros(n, dtype=np.int64)
    ci = np.zeros(n, dtype=np.int64)
    rc = _L().toist_lsap_f64(c.ctypes.data, nr, nc, ri.ctypes.data, ci.ctypes.data)
    if rc == -4:
        raise ValueError(_L().toist_last_error().decode())
    if rc < 0:
        _lib.check(rc)
    return ri, ci


def relu_bwd(dy: torch.Tensor, y: torch.Tensor, dy2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(dy + dy2) * (y > 0), bf16."""
    assert dy.is_contiguous() and y.is_contiguous() and dy.dtype == torch.bfloat16
    out = torch.empty_like(dy)
    _ck(_L().toist_relu_bwd(dy.data_ptr(), _ptr(dy2), y.data_ptr(), out.data_ptr(), dy.numel(), _stream()))
    return out


# ------------------------------------------------------------------------------------------------ criterion
def token_ce(logits, match_q, tgt_count, posmap, num_boxes, eos_coef: float, want_grad: bool):
    L, B, Q, Cc = logits.shape
    tmax = match_q.shape[-1]
    row_loss = torch.empty((L, B, Q), dtype=torch.float32, device=logits.device)
    dlogits = torch.empty_like(logits) if want_grad else None
    _ck(_L().toist_token_ce(logits.data_ptr(), match_q.data_ptr(), tgt_count.data_ptr(), posmap.data_ptr(),
                            num_boxes.data_ptr(), row_loss.data_ptr(), _ptr(dlogits), L, B, Q, Cc, tmax, eos_coef,
                            _stream()))
    return row_loss, dlogits


def box_loss(boxes, match_q, tgt_count, tgt_boxes, num_boxes, want_grad: bool):
    L, B, Q, _ = boxes.shape
    tmax = match_q.shape[-1]
    dev = boxes.device
    pl1 = torch.empty((L, B, tmax), dtype=torch.float32, device=dev)
    pgi = torch.empty((L, B, tmax), dtype=torch.float32, device=dev)
    d1 = torch.empty_like(boxes) if want_grad else None
    d2 = torch.empty_like(boxes) if want_grad else None
    _ck(_L().toist_box_loss(boxes.data_ptr(), match_q.data_ptr(), tgt_count.data_ptr(), tgt_boxes.data_ptr(),
                            num_boxes.data_ptr(), pl1.data_ptr(), pgi.data_ptr(), _ptr(d1), _ptr(d2), L, B, Q, tmax,
                            _stream()))
    return pl1, pgi, d1, d2


def cardinality(logits) -> torch.Tensor:
    L, B, Q, Cc = logits.shape
    card = torch.empty((L, B), dtype=torch.int32, device=logits.device)
    _ck(_L().toist_cardinality(logits.data_ptr(), card.data_ptr(), L, B, Q, Cc, _stream()))
    return card


def contrastive_align(pq, pt, match_q, tgt_count, tok_pos, num_boxes, temperature: float, want_grad: bool = False):
    L, B, Q, D = pq.shape
    Kt = pt.shape[1]
    tmax = match_q.shape[-1]
    img_loss = torch.empty((L, B), dtype=torch.float32, device=pq.device)
    dpq = torch.empty_like(pq) if want_grad else None
    dpt = torch.empty((L, B, Kt, D), dtype=torch.float32, device=pq.device) if want_grad else None
    _ck(_L().toist_contrastive_align(pq.data_ptr(), pt.data_ptr(), match_q.data_ptr(), tgt_count.data_ptr(),
                                     tok_pos.data_ptr(), num_boxes.data_ptr(), img_loss.data_ptr(), _ptr(dpq),
                                     _ptr(dpt), L, B, Q, Kt, D, tmax, temperature, _stream()))
    return img_loss, dpq, dpt


def criterion_reduce(row_loss, pl1, pgi, card, img_loss, tgt_count, num_boxes, flags) -> torch.Tensor:
    L, B, Q = row_loss.shape
    tmax = pl1.shape[-1]
    out = torch.zeros((5, L), dtype=torch.float32, device=row_loss.device)
    _ck(_L().toist_criterion_reduce(row_loss.data_ptr(), pl1.data_ptr(), pgi.data_ptr(), card.data_ptr(),
                                    _ptr(img_loss), tgt_count.data_ptr(), num_boxes.data_ptr(), _ptr(flags),
                                    out.data_ptr(), L, B, Q, tmax, _stream()))
    return out


def scale_layers(x: torch.Tensor, g: torch.Tensor, reduce: bool = False) -> torch.Tensor:
    """y[l] = x[l] * g[l] for x [L, ...] fp32, g [L] fp32; with `reduce` y = sum_l x[l] * g[l] (shape x.shape[1:])."""
    L = x.shape[0]
    assert x.dtype == torch.float32 and x.is_contiguous() and g.is_contiguous() and g.numel() == L
    y = torch.empty(x.shape[1:], dtype=x.dtype, device=x.device) if reduce else torch.empty_like(x)
    _ck(_L().toist_scale_layers(x.data_ptr(), g.data_ptr(), y.data_ptr(), L, x.numel() // L, 1 if reduce else 0,
                                _stream()))
    return y


def scale_layers2(x1, g1, x2, g2) -> torch.Tensor:
    L = x1.shape[0]
    assert x1.is_contiguous() and x2.is_contiguous() and g1.is_contiguous() and g2.is_contiguous()
    y = torch.empty_like(x1)
    _ck(_L().toist_scale_layers2(x1.data_ptr(), g1.data_ptr(), x2.data_ptr(), g2.data_ptr(), y.data_ptr(), L,
                                 x1.numel() // L, _stream()))
    return y


# ------------------------------------------------------------------------------------------------ mask branch
def attn_map_fwd(q: torch.Tensor, k: torch.Tensor, key_mask_u8: Optional[torch.Tensor], nhead: int) -> torch.Tensor:
    """MHAttentionMap core (models/segmentation.py:262-273): softmax_k(q . k * dh**-0.5) per head, no value product.
    q [Sq, B, E], k [Sk, B, E] bf16 -> probs bf16 [B, H, Sq, ld]."""
    sq, b, e = q.shape
    sk = k.shape[0]
    d = e // nhead
    ld = _ld8(sk)
    scores = torch.empty((b, nhead, sq, ld), dtype=torch.float32, device=q.device)
    gemm(GEMM_FWD, t4(q, (d, sq, nhead, b), (1, q.stride(0), d, q.stride(1))),
         t4(k, (d, sk, nhead, b), (1, k.stride(0), d, k.stride(1))), scores, ext=(sq, nhead, b), tile=(128, 1, 1),
         n_cols=sk, out_strides=(ld, sq * ld, nhead * sq * ld), k_per_tap=d, b_batched=True, alpha=float(d) ** -0.5)
    probs = torch.empty((b, nhead, sq, ld), dtype=torch.bfloat16, device=q.device)
    _ck(_L().toist_attn_softmax_fwd(scores.data_ptr(), _ptr(key_mask_u8), probs.data_ptr(), None, b * nhead * sq, sk, ld,
                                    ld, nhead * sq, 0.0, None, 0, _stream()))
    return probs


def attn_map_bwd(dprobs: torch.Tensor, q: torch.Tensor, k: torch.Tensor, probs: torch.Tensor, nhead: int,
                 dq: torch.Tensor, dk: torch.Tensor) -> None:
    """dprobs fp32 [B, H, Sq, ld] -> dq, dk (bf16, same indexing as q, k)."""
    sq, b, e = q.shape
    sk = k.shape[0]
    d = e // nhead
    ld = probs.shape[-1]
    sP = (1, ld, sq * ld, nhead * sq * ld)
    ds = torch.empty((b, nhead, sq, ld), dtype=torch.bfloat16, device=q.device)
    _ck(_L().toist_attn_softmax_bwd(dprobs.data_ptr(), probs.data_ptr(), ds.data_ptr(), b * nhead * sq, sk, ld, ld,
                                    float(d) ** -0.5, 0.0, None, 0, _stream()))
    gemm(GEMM_DGRAD, t4(ds, (sk, sq, nhead, b), sP), t4(k, (d, sk, nhead, b), (1, k.stride(0), d, k.stride(1))), dq,
         ext=(sq, nhead, b), tile=(128, 1, 1), n_cols=d, out_strides=(dq.stride(0), d, dq.stride(1)), k_per_tap=sk,
         b_batched=True)
    gemm(GEMM_WGRAD, t4(ds, (sk, sq, nhead, b), sP), t4(q, (d, sq, nhead, b), (1, q.stride(0), d, q.stride(1))), dk,
         ext=(sq, 1, 1), tile=(64, 1, 1), n_cols=d, m_rows=sk, out_strides=(dk.stride(0), d, dk.stride(1)),
         batch=(nhead, b))


def mask_input(src_proj: torch.Tensor, probs: torch.Tensor, B: int, Q: int, h: int, w: int) -> torch.Tensor:
    """src_proj bf16 [hw, B, E], probs bf16 [B, NH, Q, ld] -> x0 bf16 NHWC [B*Q, h, w, E + NH]."""
    hw, _, E = src_proj.shape
    NH, ld = probs.shape[1], probs.shape[-1]
    assert src_proj.is_contiguous() and probs.is_contiguous() and hw == h * w
    x0 = torch.empty((B * Q, h, w, E + NH), dtype=torch.bfloat16, device=src_proj.device)
    _ck(_L().toist_mask_input(src_proj.data_ptr(), probs.data_ptr(), x0.data_ptr(), B, Q, hw, E, NH, ld, _stream()))
    return x0


def mask_input_bwd(dx0: torch.Tensor, B: int, Q: int, E: int, NH: int, ld: int, want_src: bool):
    n, h, w, c = dx0.shape
    hw = h * w
    dsrc = torch.empty((hw, B, E), dtype=torch.bfloat16, device=dx0.device) if want_src else None
    dattn = torch.zeros((B, NH, Q, ld), dtype=torch.float32, device=dx0.device)
    _ck(_L().toist_mask_input_bwd(dx0.data_ptr(), _ptr(dsrc), dattn.data_ptr(), B, Q, hw, E, NH, ld, _stream()))
    return dsrc, dattn


def groupnorm_relu_fwd(z: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 8, eps: float = 1e-5):
    """z NHWC bf16 [N, H, W, C] -> (a, mean [N, G], rstd [N, G])."""
    n, h, w, c = z.shape
    assert z.is_contiguous() and z.dtype == torch.bfloat16
    a = torch.empty_like(z)
    mean = torch.empty((n, groups), dtype=torch.float32, device=z.device)
    rstd = torch.empty((n, groups), dtype=torch.float32, device=z.device)
    _ck(_L().toist_groupnorm_relu_fwd(z.data_ptr(), gamma.data_ptr(), beta.data_ptr(), a.data_ptr(), mean.data_ptr(),
                                      rstd.data_ptr(), n, h * w, c, groups, eps, _stream()), 2)
    return a, mean, rstd


def groupnorm_relu_bwd(da: torch.Tensor, z: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor,
                       beta: torch.Tensor, dgamma: Optional[torch.Tensor], dbeta: Optional[torch.Tensor],
                       groups: int = 8) -> torch.Tensor:
    n, h, w, c = z.shape
    assert da.is_contiguous() and da.shape == z.shape
    dz = torch.empty_like(z)
    scratch = torch.empty((2, n, groups), dtype=torch.float32, device=z.device)
    _ck(_L().toist_groupnorm_relu_bwd(da.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                      beta.data_ptr(), dz.data_ptr(), _ptr(dgamma), _ptr(dbeta), scratch.data_ptr(), n,
                                      h * w, c, groups, _stream()), 2)
    return dz


def upsample_add(xs: torch.Tensor, fpn: torch.Tensor, Q: int) -> torch.Tensor:
    """fpn[n // Q] + nearest_upsample(xs) : xs [N, h, w, C], fpn [N / Q, H, W, C] bf16 NHWC."""
    n, h, w, c = xs.shape
    nb, H, W, c2 = fpn.shape
    assert c == c2 and nb * Q == n and xs.is_contiguous() and fpn.is_contiguous()
    out = torch.empty((n, H, W, c), dtype=torch.bfloat16, device=xs.device)
    _ck(_L().toist_upsample_add(xs.data_ptr(), fpn.data_ptr(), out.data_ptr(), n, Q, H, W, h, w, c, _stream()))
    return out


def upsample_add_bwd(dout: torch.Tensor, in_hw: Tuple[int, int], Q: int, want_fpn: bool):
    n, H, W, c = dout.shape
    h, w = in_hw
    dxs = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=dout.device)
    dfpn = torch.empty((n // Q, H, W, c), dtype=torch.bfloat16, device=dout.device) if want_fpn else None
    _ck(_L().toist_upsample_add_bwd(dout.data_ptr(), dxs.data_ptr(), _ptr(dfpn), n, Q, H, W, h, w, c, _stream()), 2)
    return dxs, dfpn


def mask_loss_fwd(pred_masks, tgt_masks_u8, match_q, tgt_count, num_boxes):
    B, Q, mh, mw = pred_masks.shape
    _, tmax, th, tw = tgt_masks_u8.shape
    sums = torch.empty((B, tmax, 4), dtype=torch.float32, device=pred_masks.device)
    out = torch.empty(2, dtype=torch.float32, device=pred_masks.device)
    _ck(_L().toist_mask_loss_fwd(pred_masks.data_ptr(), tgt_masks_u8.data_ptr(), match_q.data_ptr(), tgt_count.data_ptr(),
                                 num_boxes.data_ptr(), sums.data_ptr(), out.data_ptr(), B, Q, tmax, mh, mw, th, tw,
                                 _stream()), 2)
    return out, sums


def mask_loss_bwd(pred_masks, tgt_masks_u8, match_q, tgt_count, sums, num_boxes, gout):
    B, Q, mh, mw = pred_masks.shape
    _, tmax, th, tw = tgt_masks_u8.shape
    dpred = torch.empty_like(pred_masks)
    _ck(_L().toist_mask_loss_bwd(pred_masks.data_ptr(), tgt_masks_u8.data_ptr(), match_q.data_ptr(), tgt_count.data_ptr(),
                                 sums.data_ptr(), num_boxes.data_ptr(), gout.data_ptr(), dpred.data_ptr(), B, Q, tmax,
                                 mh, mw, th, tw, _stream()))
    return dpred


# ------------------------------------------------------------------------------------------------ distillation
def softkd_fwd(logits_n, logits_s, boxes_n, boxes_s, match_n, match_s, tgt_count, flags):
    """Soft-KD loss of every decoder layer (models/mdetr.py:543-599): returns (loss [L], saved workspace)."""
    L, B, Q, C = logits_s.shape
    tmax = match_s.shape[-1]
    dev = logits_s.device
    f32, i32 = torch.float32, torch.int32
    bi_n = torch.empty((L, B, Q, 2), dtype=f32, device=dev)
    bi_s = torch.empty((L, B, Q, 2), dtype=f32, device=dev)
    fp_n = torch.empty((L, B, Q), dtype=i32, device=dev)
    fp_s = torch.empty((L, B, Q), dtype=i32, device=dev)
    n_fp = torch.empty((2, L * B), dtype=i32, device=dev)
    cost = torch.empty((L, B, Q, Q), dtype=f32, device=dev)
    col = torch.empty((L, B, Q), dtype=i32, device=dev)
    pair = torch.empty((L, B, Q), dtype=i32, device=dev)
    loss = torch.empty((L,), dtype=f32, device=dev)
    for t in (logits_n, logits_s, boxes_n, boxes_s, match_n, match_s):
        assert t.is_contiguous()
    _ck(_L().toist_softkd_fwd(logits_n.data_ptr(), logits_s.data_ptr(), boxes_n.data_ptr(), boxes_s.data_ptr(),
                              match_n.data_ptr(), match_s.data_ptr(), tgt_count.data_ptr(), bi_n.data_ptr(),
                              bi_s.data_ptr(), fp_n.data_ptr(), fp_s.data_ptr(), n_fp.data_ptr(), cost.data_ptr(),
                              col.data_ptr(), pair.data_ptr(), flags.data_ptr(), loss.data_ptr(), L, B, Q, C, tmax,
                              _stream()), 6)
    return loss, (bi_n, bi_s, pair, n_fp, cost, col, fp_n, fp_s)


def softkd_bwd(logits_s, bi_n, bi_s, pair, tgt_count, n_fp, gout, tmax: int):
    L, B, Q, C = logits_s.shape
    d = torch.empty_like(logits_s)
    _ck(_L().toist_softkd_bwd(logits_s.data_ptr(), bi_n.data_ptr(), bi_s.data_ptr(), pair.data_ptr(),
                              tgt_count.data_ptr(), n_fp.data_ptr(), gout.data_ptr(), d.data_ptr(), L, B, Q, C, tmax,
                              _stream()))
    return d


def lsap_batched(cost: torch.Tensor, n_rows: torch.Tensor, n_cols: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """cost fp32 [P, R, C]; n_rows / n_cols int32 [P]  ->  col_of_row int32 [P, R] (-1 = unassigned)."""
    P, R, Cc = cost.shape
    assert cost.is_contiguous() and cost.dtype == torch.float32
    out = torch.empty((P, R), dtype=torch.int32, device=cost.device)
    _ck(_L().toist_lsap_batched(cost.data_ptr(), n_rows.data_ptr(), n_cols.data_ptr(), out.data_ptr(), flags.data_ptr(),
                                P, R, Cc, _stream()))
    return out


def kmeans(x: torch.Tensor, centers: torch.Tensor, tol: float = 1e-4, max_iter: int = 10000):
    """In-place Lloyd iterations on `centers` [K, D] (models/kmeans.py:21-96); returns (choice int32 [N], iters [1])."""
    n, d = x.shape
    k = centers.shape[0]
    assert x.is_contiguous() and centers.is_contiguous() and x.dtype == centers.dtype == torch.float32
    choice = torch.empty((n,), dtype=torch.int32, device=x.device)
    iters = torch.empty((1,), dtype=torch.int32, device=x.device)
    xt = torch.empty((d, n), dtype=torch.float32, device=x.device)  # transposed bank (coalesced assignment step)
    _ck(_L().toist_kmeans(x.data_ptr(), centers.data_ptr(), choice.data_ptr(), iters.data_ptr(), xt.data_ptr(), n, d, k,
                          float(tol), int(max_iter), _stream()))
    return choice, iters


def kmeans_batched(banks: torch.Tensor, task_of: torch.Tensor, centers: torch.Tensor, query: torch.Tensor,
                   tol: float = 1e-4, max_iter: int = 10000) -> torch.Tensor:
    """Independent k-means problems in one launch: problem p clusters banks[task_of[p]] ([N, D]) from centers[p] [K, D]
    (updated in place).  Returns the index [P] of the final centre nearest to query[p] [D]."""
    t, n, d = banks.shape
    p, k, _ = centers.shape
    assert banks.is_contiguous() and centers.is_contiguous() and query.is_contiguous() and query.shape == (p, d)
    assert task_of.dtype == torch.int32 and task_of.numel() == p
    dev = banks.device
    choice = torch.empty((p, n), dtype=torch.int32, device=dev)
    xt = torch.empty((p, d, n), dtype=torch.float32, device=dev)
    qc = torch.empty((p,), dtype=torch.int32, device=dev)
    _ck(_L().toist_kmeans_batched(banks.data_ptr(), task_of.data_ptr(), centers.data_ptr(), choice.data_ptr(), None,
                                  xt.data_ptr(), query.data_ptr(), qc.data_ptr(), p, n, d, k, float(tol), int(max_iter),
                                  _stream()))
    return qc


def kmeans_predict(x: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    m, d = x.shape
    choice = torch.empty((m,), dtype=torch.int32, device=x.device)
    _ck(_L().toist_kmeans_predict(x.data_ptr(), centers.data_ptr(), choice.data_ptr(), m, d, centers.shape[0], _stream()))
    return choice


def token_wsum(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x fp32 [T, B, D], w fp32 [B, T] -> out[b] = sum_t w[b, t] x[t, b]."""
    T, B, D = x.shape
    assert x.is_contiguous() and w.is_contiguous() and w.shape == (B, T) and w.dtype == torch.float32
    out = torch.empty((B, D), dtype=torch.float32, device=x.device)
    _ck(_L().toist_token_wsum(x.data_ptr(), w.data_ptr(), out.data_ptr(), T, B, D, _stream()))
    return out


def token_wsum_bwd(dout: torch.Tensor, w: torch.Tensor, T: int) -> torch.Tensor:
    B, D = dout.shape
    dx = torch.empty((T, B, D), dtype=torch.float32, device=dout.device)
    _ck(_L().toist_token_wsum_bwd(dout.data_ptr(), w.data_ptr(), dx.data_ptr(), T, B, D, _stream()))
    return dx


def token_fill(x: torch.Tensor, sel_u8: torch.Tensor, feat: Optional[torch.Tensor]) -> torch.Tensor:
    """In place: x[t, b] = feat[b] (zeros when feat is None) for the selected tokens; x fp32 [T, B, D] contiguous."""
    T, B, D = x.shape
    assert x.is_contiguous() and sel_u8.is_contiguous() and sel_u8.shape == (B, T)
    assert feat is None or (feat.is_contiguous() and feat.shape == (B, D))
    _ck(_L().toist_token_fill(x.data_ptr(), sel_u8.data_ptr(), _ptr(feat), T, B, D, _stream()))
    return x


def mse_rows(a: torch.Tensor, b: torch.Tensor, use_u8: torch.Tensor, want_grad: bool):
    M, D = a.shape
    loss = torch.empty((1,), dtype=torch.float32, device=a.device)
    da = torch.empty_like(a) if want_grad else None
    _ck(_L().toist_mse_rows(a.data_ptr(), b.data_ptr(), use_u8.data_ptr(), loss.data_ptr(), _ptr(da), M, D, _stream()))
    return loss, da


def cdist_l1(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    _ck(_L().toist_cdist_l1(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.shape[0], b.shape[0], a.shape[1], _stream()))
    return out


# ------------------------------------------------------------------------------------------------ path ends (csrc/io.cu)
def _image_table(images: Sequence[torch.Tensor], hw: Sequence[Tuple[int, int]], dev):
    """Device arrays (pointers [B] uint64, sizes [B, 2] int32) describing a list of per-image device tensors."""
    import numpy as np

    from .util.misc import h2d

    ptrs = torch.from_numpy(np.asarray([t.data_ptr() for t in images], dtype=np.int64))
    sizes = torch.tensor([[int(h), int(w)] for h, w in hw], dtype=torch.int32)
    return h2d(ptrs, dev), h2d(sizes, dev)


def pad_normalize_u8(images: Sequence[torch.Tensor], mean, std, height: int, width: int):
    """uint8 HWC images (CUDA, contiguous, 3 channels) -> (fp32 [B, 3, H, W] normalised + zero padded, uint8 mask
    [B, H, W] with 1 = padding): ToTensor + Normalize + NestedTensor padding in one launch."""
    dev = images[0].device
    for t in images:
        assert t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous() and t.dim() == 3 and t.shape[2] == 3
    B = len(images)
    ptrs, sizes = _image_table(images, [(t.shape[0], t.shape[1]) for t in images], dev)
    out = torch.empty((B, 3, height, width), dtype=torch.float32, device=dev)
    mask = torch.empty((B, height, width), dtype=torch.uint8, device=dev)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    _ck(_L().toist_pad_normalize_u8(ptrs.data_ptr(), sizes.data_ptr(), out.data_ptr(), mask.data_ptr(), B, height, width,
                                    C.cast(m, C.c_void_p), C.cast(s, C.c_void_p), _stream()))
    return out, mask


def pad_batch_f32(images: Sequence[torch.Tensor], height: int, width: int):
    """fp32 CHW images (CUDA, contiguous) -> (zero padded [B, C, H, W], uint8 mask [B, H, W]) in one launch."""
    dev = images[0].device
    c = images[0].shape[0]
    for t in images:
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 3 and t.shape[0] == c
    B = len(images)
    ptrs, sizes = _image_table(images, [(t.shape[1], t.shape[2]) for t in images], dev)
    out = torch.empty((B, c, height, width), dtype=torch.float32, device=dev)
    mask = torch.empty((B, height, width), dtype=torch.uint8, device=dev)
    _ck(_L().toist_pad_batch_f32(ptrs.data_ptr(), sizes.data_ptr(), out.data_ptr(), mask.data_ptr(), B, c, height, width,
                                 _stream()))
    return out, mask


def postprocess_boxes(logits: torch.Tensor, boxes: torch.Tensor, target_sizes: torch.Tensor,
                      is_final: Optional[torch.Tensor] = None):
    """models/postprocessors.py:15-58 -> (scores [B, Q], labels [B, Q] int64, boxes xyxy absolute [B, Q, 4], scores_refexp)."""
    B, Q, Cc = logits.shape
    logits, boxes = logits.float().contiguous(), boxes.float().contiguous()
    ts = target_sizes.contiguous()
    assert ts.shape == (B, 2) and ts.is_cuda
    if ts.dtype not in (torch.int64, torch.float32):
        ts = ts.to(torch.float32)
    dev = logits.device
    scores = torch.empty((B, Q), dtype=torch.float32, device=dev)
    labels = torch.empty((B, Q), dtype=torch.int64, device=dev)
    out = torch.empty((B, Q, 4), dtype=torch.float32, device=dev)
    fin = ref = None
    if is_final is not None:
        fin = is_final.float().contiguous().view(B, Q)
        ref = torch.empty((B, Q), dtype=torch.float32, device=dev)
    _ck(_L().toist_postprocess_boxes(logits.data_ptr(), boxes.data_ptr(), ts.data_ptr() if ts.dtype == torch.float32 else None,
                                     ts.data_ptr() if ts.dtype == torch.int64 else None, _ptr(fin), scores.data_ptr(),
                                     labels.data_ptr(), out.data_ptr(), _ptr(ref), B, Q, Cc, _stream()))
    return scores, labels, out, ref


def postprocess_masks(pred: torch.Tensor, stage1_hw, crop_hw, out_hw, threshold: float) -> torch.Tensor:
    """pred [Q, h, w] fp32 of one image -> bool [Q, out_h, out_w] (models/postprocessors.py:79-107, fused)."""
    assert pred.is_cuda and pred.dtype == torch.float32 and pred.is_contiguous() and pred.dim() == 3
    Q, hm, wm = pred.shape
    out = torch.empty((Q, int(out_hw[0]), int(out_hw[1])), dtype=torch.uint8, device=pred.device)
    _ck(_L().toist_postprocess_masks(pred.data_ptr(), out.data_ptr(), Q, hm, wm, int(stage1_hw[0]), int(stage1_hw[1]),
                                     int(crop_hw[0]), int(crop_hw[1]), int(out_hw[0]), int(out_hw[1]), float(threshold),
                                     _stream()))
    return out.view(torch.bool)
