"""Tensor-level wrappers over the C ABI (no autograd here; see functional.py for the autograd.Functions).

Every function launches hand-written sm_100a kernels from libtoist_b200.so on torch's current CUDA stream.
torch is used only to own device memory.  Layout conventions:
  * activations of the convolutional trunk are NHWC ("channels last") bf16,
  * sequence activations are [rows, features] bf16 with features contiguous,
  * weights are bf16 [out_features, in_features] (conv: [Cout, kh, kw, Cin]), master copies stay fp32 nn.Parameters.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, BF16, F32, GEMM_DGRAD, GEMM_FWD, GEMM_WGRAD, GemmDesc,
                   Tap, Tensor4)

_launch_count = 0


def launches() -> int:
    """Number of toist_b200 kernel launches issued by this process (bench.py reports the per-step delta)."""
    return _launch_count


def _count(n: int = 1) -> None:
    global _launch_count
    _launch_count += n


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
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx[M,K] = (dy[M,N] @ w[N,K] + res) * (mask > 0)."""
    M, N = dy.shape
    K = w.shape[1]
    assert w.shape[0] == N and dy.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, K), dtype=out_dtype, device=dy.device)
    gemm(GEMM_DGRAD, t4(dy, (N, M, 1, 1), (1, dy.stride(0), 0, 0)), t4(w, (K, N, 1, 1), (1, w.stride(0), 0, 0)), out,
         ext=(M, 1, 1), tile=(128, 1, 1), n_cols=K, out_strides=(out.stride(0), 0, 0), k_per_tap=N, res=res,
         mask=mask)
    return out


def _wgrad_splits(m_rows: int, n_cols: int, n_taps: int, pixel_tiles: int, batch: int = 1) -> int:
    base = -(-m_rows // 128) * -(-n_cols // 128) * n_taps * batch
    want = max(1, 296 // max(base, 1))
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
             scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[N,Ho,Wo,Cout] = act((conv(x[N,H,W,Cin], w[Cout,kh,kw,Cin]) * scale + shift) + res), all NHWC bf16."""
    n, h, wd, cin = x.shape
    cout, kh, kw, cin2 = w.shape
    assert cin == cin2 and w.is_contiguous()
    ho, wo = conv_out_size(h, kh, stride, pad), conv_out_size(wd, kw, stride, pad)
    out = torch.empty((n, ho, wo, cout), dtype=out_dtype, device=x.device)
    taps = [(kx - pad, ky - pad, 0, (ky * kw + kx) * cin) for ky in range(kh) for kx in range(kw)]
    tile = pick_tile(wo, ho, n, 128, 128 // stride if stride > 1 else 128)
    gemm(GEMM_FWD, _nhwc_t4(x), t4(w, (kh * kw * cin, cout, 1, 1), (1, kh * kw * cin, 0, 0)), out,
         ext=(wo, ho, n), tile=tile, n_cols=cout, out_strides=(cout, wo * cout, ho * wo * cout), k_per_tap=cin,
         taps=taps, stride=(stride, stride), col_scale=scale, col_shift=shift, res=res, act=act)
    return out


def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, in_hw: Tuple[int, int], *, stride: int = 1, pad: int = 0,
               mask: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx[N,H,W,Cin] = (conv_transpose(dy[N,Ho,Wo,Cout], w[Cout,kh,kw,Cin]) + res) * (mask > 0)."""
    n, ho, wo, cout = dy.shape
    cout2, kh, kw, cin = w.shape
    assert cout == cout2 and w.is_contiguous()
    h, wd = in_hw
    dx = torch.empty((n, h, wd, cin), dtype=torch.bfloat16, device=dy.device)
    a = _nhwc_t4(dy)
    b = t4(w, (kh * kw * cin, cout, 1, 1), (1, kh * kw * cin, 0, 0))
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
    n, ho, wo, cout = dy.shape
    n2, h, wd, cin = x.shape
    cout2, kh, kw, cin2 = dw.shape
    assert n == n2 and cout == cout2 and cin == cin2 and dw.dtype == torch.float32 and dw.is_contiguous()
    taps = [(kx - pad, ky - pad, 0, (ky * kw + kx) * cin) for ky in range(kh) for kx in range(kw)]
    tile = pick_tile(wo, ho, n, 64, 64)
    ptiles = -(-wo // tile[0]) * -(-ho // tile[1]) * -(-n // tile[2])
    splits = _wgrad_splits(cout, cin, len(taps), ptiles) if accumulate else 1
    gemm(GEMM_WGRAD, _nhwc_t4(dy), _nhwc_t4(x), dw, ext=(wo, ho, n), tile=tile, n_cols=cin, m_rows=cout,
         out_strides=(kh * kw * cin, 0, 0), taps=taps, stride=(stride, stride), splits=splits, row_scale=row_scale,
         accumulate=accumulate)
    return dw
