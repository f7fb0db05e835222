// Row-wise normalisations and softmax (one warp per row, warp-shuffle reductions, fp32 statistics):
// LayerNorm fwd/bwd (models/transformer.py:279-280,346-349,484; RoBERTa LayerNorms), L2 normalise fwd/bwd
// (models/mdetr.py:430-433), key-padding-masked attention softmax fwd/bwd (nn.MultiheadAttention,
// models/transformer.py:298,378,394), sine position embedding (models/position_encoding.py:30-49),
// RoBERTa embedding gather / scatter.
#include <math.h>

#include <initializer_list>

#include "common.cuh"
#include "host_util.h"

namespace toist {

constexpr int kMaxPerLane = 32;  // rows up to 1024 wide

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p, long long i) {
  if constexpr (sizeof(T) == 2) return __bfloat162float(p[i]);
  else return p[i];
}
template <typename T>
__device__ __forceinline__ void st_from_float(T* p, long long i, float v) {
  if constexpr (sizeof(T) == 2) p[i] = __float2bfloat16_rn(v);
  else p[i] = v;
}

// ------------------------------------------------------------------------------------------------ LayerNorm
template <typename TIn>
__global__ void layernorm_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ y16,
                                     float* __restrict__ y32, float* __restrict__ mean, float* __restrict__ rstd,
                                     long long rows, int N, float eps) {
  pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const TIn* xr = x + row * N;
  float v[kMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxPerLane; ++k) {
    const int c = lane + 32 * k;
    v[k] = c < N ? ld_as_float(xr, c) : 0.f;
    s += v[k];
  }
  const float mu = warp_sum(s) / (float)N;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxPerLane; ++k) {
    const int c = lane + 32 * k;
    const float d = c < N ? v[k] - mu : 0.f;
    sq += d * d;
  }
  const float rs = rsqrtf(warp_sum(sq) / (float)N + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
#pragma unroll
  for (int k = 0; k < kMaxPerLane; ++k) {
    const int c = lane + 32 * k;
    if (c < N) {
      const float o = (v[k] - mu) * rs * gamma[c] + beta[c];
      if (y16) y16[row * N + c] = __float2bfloat16_rn(o);
      if (y32) y32[row * N + c] = o;
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma += dy * xhat;  dbeta += dy
template <typename TX, typename TDy, typename TDx>
__global__ void layernorm_bwd_kernel(const TDy* __restrict__ dy, const TDy* __restrict__ dy2, const TX* __restrict__ x,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     const float* __restrict__ gamma, TDx* __restrict__ dx, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, long long rows, int N) {
  pdl_prologue();
  extern __shared__ float acc[];  // [2][N]
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const float mu = mean[row], rs = rstd[row];
    float xh[kMaxPerLane], g[kMaxPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int c = lane + 32 * k;
      if (c < N) {
        float d = ld_as_float(dy, row * N + c);
        if (dy2) d += ld_as_float(dy2, row * N + c);
        xh[k] = (ld_as_float(x, row * N + c) - mu) * rs;
        g[k] = d * gamma[c];
        s1 += g[k];
        s2 += g[k] * xh[k];
        if (dgamma) {
          atomicAdd(&acc[c], d * xh[k]);
          atomicAdd(&acc[N + c], d);
        }
      } else {
        xh[k] = 0.f;
        g[k] = 0.f;
      }
    }
    s1 = warp_sum(s1) / (float)N;
    s2 = warp_sum(s2) / (float)N;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int c = lane + 32 * k;
      if (c < N) st_from_float(dx, row * N + c, rs * (g[k] - s1 - xh[k] * s2));
    }
  }
  __syncthreads();
  if (dgamma)
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      atomicAdd(dgamma + i, acc[i]);
      atomicAdd(dbeta + i, acc[N + i]);
    }
}


// ------------------------------------------------------------------------------------------------ LayerNorm, bf16 fast path
// Rows of N = 256 * K bf16 elements (the transformer's 256, RoBERTa's 768): lane l owns the 8 contiguous columns
// 256 * k + 8 * l .. + 7 of every 256-wide slice, so a row is K 16-byte loads per lane (full 512-byte warp
// transactions), the statistics are two shuffle reductions and nothing is guarded.  The backward keeps the
// dgamma / dbeta partial sums of the lane's own columns in registers across the rows of a grid-stride loop; the CTA
// reduces them once through shared memory and issues one global atomic per column (the generic kernel above spends
// two shared-memory atomics per ELEMENT).
template <int K>
__global__ void __launch_bounds__(128)
layernorm_fwd_vec_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                         const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
                         float* __restrict__ rstd, long long rows, float eps) {
  constexpr int N = 256 * K;
  const int lane = threadIdx.x & 31;
  float ga[K][8], be[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k) {  // parameters do not depend on the previous kernel: load before the PDL wait
    const float4* g4 = reinterpret_cast<const float4*>(gamma + 256 * k + 8 * lane);
    const float4* b4 = reinterpret_cast<const float4*>(beta + 256 * k + 8 * lane);
    const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
    ga[k][0] = g0.x; ga[k][1] = g0.y; ga[k][2] = g0.z; ga[k][3] = g0.w;
    ga[k][4] = g1.x; ga[k][5] = g1.y; ga[k][6] = g1.z; ga[k][7] = g1.w;
    be[k][0] = b0.x; be[k][1] = b0.y; be[k][2] = b0.z; be[k][3] = b0.w;
    be[k][4] = b1.x; be[k][5] = b1.y; be[k][6] = b1.z; be[k][7] = b1.w;
  }
  pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + row * N;
  float v[K][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + 256 * k + 8 * lane);
    const uint32_t* pu = &u.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16(pu[j]);
      v[k][2 * j] = f.x;
      v[k][2 * j + 1] = f.y;
      s += f.x + f.y;
    }
  }
  const float mu = warp_sum(s) * (1.f / N);
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[k][j] - mu;
      sq += d * d;
    }
  const float rs = rsqrtf(warp_sum(sq) * (1.f / N) + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mu) * rs * ga[k][j] + be[k][j];
    uint4 u;
    u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
    *reinterpret_cast<uint4*>(y + row * N + 256 * k + 8 * lane) = u;
  }
}

template <int K>
__global__ void __launch_bounds__(128)
layernorm_bwd_vec_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                         const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                         __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         long long rows, __nv_bfloat16* __restrict__ dx_drop,
                         const unsigned long long* __restrict__ seed, uint32_t site, uint32_t thr16, float scale) {
  // dx_drop (nullable) = dropout-backward of dx with the decisions of site `site` (the gradient w.r.t. the linear
  // layer's output behind a `res + dropout(y)` branch), computed from the bf16-rounded dx like the separate kernel did
  constexpr int N = 256 * K;
  __shared__ float red[4][2][N];  // per warp: [dgamma | dbeta] partials of the CTA's rows
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float ga[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float4* g4 = reinterpret_cast<const float4*>(gamma + 256 * k + 8 * lane);
    const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1);
    ga[k][0] = g0.x; ga[k][1] = g0.y; ga[k][2] = g0.z; ga[k][3] = g0.w;
    ga[k][4] = g1.x; ga[k][5] = g1.y; ga[k][6] = g1.z; ga[k][7] = g1.w;
  }
  pdl_prologue();
  float ag[K][8], ab[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[k][j] = ab[k][j] = 0.f;
  for (long long row = (long long)blockIdx.x * wpb + warp; row < rows; row += (long long)gridDim.x * wpb) {
    const float mu = mean[row], rs = rstd[row];
    float xh[K][8], g[K][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const long long off = row * N + 256 * k + 8 * lane;
      const uint4 ud = *reinterpret_cast<const uint4*>(dy + off);
      const uint4 ux = *reinterpret_cast<const uint4*>(x + off);
      uint4 u2 = make_uint4(0u, 0u, 0u, 0u);
      if (dy2 != nullptr) u2 = *reinterpret_cast<const uint4*>(dy2 + off);
      const uint32_t* pd = &ud.x;
      const uint32_t* px = &ux.x;
      const uint32_t* p2 = &u2.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fd = unpack_bf16(pd[j]), fx = unpack_bf16(px[j]), f2 = unpack_bf16(p2[j]);
        const float d0 = fd.x + f2.x, d1 = fd.y + f2.y;
        const float h0 = (fx.x - mu) * rs, h1 = (fx.y - mu) * rs;
        xh[k][2 * j] = h0; xh[k][2 * j + 1] = h1;
        g[k][2 * j] = d0 * ga[k][2 * j]; g[k][2 * j + 1] = d1 * ga[k][2 * j + 1];
        s1 += g[k][2 * j] + g[k][2 * j + 1];
        s2 += g[k][2 * j] * h0 + g[k][2 * j + 1] * h1;
        ag[k][2 * j] += d0 * h0; ag[k][2 * j + 1] += d1 * h1;
        ab[k][2 * j] += d0; ab[k][2 * j + 1] += d1;
      }
    }
    s1 = warp_sum(s1) * (1.f / N);
    s2 = warp_sum(s2) * (1.f / N);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rs * (g[k][j] - s1 - xh[k][j] * s2);
      uint4 u;
      u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
      const long long off = row * N + 256 * k + 8 * lane;
      *reinterpret_cast<uint4*>(dx + off) = u;
      if (dx_drop != nullptr) {
        const uint64_t key = dropout_key(seed, site);
        const uint32_t* pu = &u.x;
        uint32_t t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w = dropout_word((uint64_t)(off >> 3) * 4 + j, key);
          const float2 f = unpack_bf16(pu[j]);
          t[j] = pack_bf16((w & 0xffffu) >= thr16 ? f.x * scale : 0.f, (w >> 16) >= thr16 ? f.y * scale : 0.f);
        }
        *reinterpret_cast<uint4*>(dx_drop + off) = make_uint4(t[0], t[1], t[2], t[3]);
      }
    }
  }
  if (dgamma == nullptr) return;
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[warp][0][256 * k + 8 * lane + j] = ag[k][j];
      red[warp][1][256 * k + 8 * lane + j] = ab[k][j];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
    const int which = i / N, c = i - which * N;
    float t = 0.f;
    for (int w = 0; w < wpb; ++w) t += red[w][which][c];
    atomicAdd((which == 0 ? dgamma : dbeta) + c, t);
  }
}

// ------------------------------------------------------------------------------------------------ fused residual branch
// The residual branches of the transformer layers end in  s = res + dropout(y);  out = LayerNorm(s)  and are followed
// by  out + pos  (the positional / query embedding added before the next attention).  Unfused that is three launches
// (dropout + residual, LayerNorm, add) per branch in a chain of ~4 us kernels; here it is one: every value is rounded
// to bf16 exactly where the separate kernels rounded it, so the results are bit-identical to the unfused path.
//   x [rows, N] bf16 (the linear layer's output); res (nullable) bf16; seed (nullable): dropout on x with the paired
//   16-bit decisions of dropout_bf16_vec_kernel (same element indexing: index / 8 = row * N/8 + 32 k + lane);
//   sum_out (nullable) = bf16(res + dropout(x)), the LayerNorm input kept for the backward; y = LayerNorm;
//   y_add (nullable) = bf16(y + add).
template <int K>
__global__ void __launch_bounds__(128)
layernorm_fused_fwd_vec_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                               const __nv_bfloat16* __restrict__ add, const float* __restrict__ gamma,
                               const float* __restrict__ beta, __nv_bfloat16* __restrict__ sum_out,
                               __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y_add,
                               float* __restrict__ mean, float* __restrict__ rstd, long long rows, float eps,
                               const unsigned long long* __restrict__ seed, uint32_t site, uint32_t thr16, float scale) {
  constexpr int N = 256 * K;
  const int lane = threadIdx.x & 31;
  float ga[K][8], be[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float4* g4 = reinterpret_cast<const float4*>(gamma + 256 * k + 8 * lane);
    const float4* b4 = reinterpret_cast<const float4*>(beta + 256 * k + 8 * lane);
    const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
    ga[k][0] = g0.x; ga[k][1] = g0.y; ga[k][2] = g0.z; ga[k][3] = g0.w;
    ga[k][4] = g1.x; ga[k][5] = g1.y; ga[k][6] = g1.z; ga[k][7] = g1.w;
    be[k][0] = b0.x; be[k][1] = b0.y; be[k][2] = b0.z; be[k][3] = b0.w;
    be[k][4] = b1.x; be[k][5] = b1.y; be[k][6] = b1.z; be[k][7] = b1.w;
  }
  pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  uint64_t key = 0;
  if (seed != nullptr) key = dropout_key(seed, site);
  float v[K][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const long long off = row * N + 256 * k + 8 * lane;
    const uint4 ux = *reinterpret_cast<const uint4*>(x + off);
    uint4 ur = make_uint4(0u, 0u, 0u, 0u);
    if (res != nullptr) ur = *reinterpret_cast<const uint4*>(res + off);
    const uint32_t* px = &ux.x;
    const uint32_t* pr = &ur.x;
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = unpack_bf16(px[j]);
      const float2 r = unpack_bf16(pr[j]);
      if (seed != nullptr) {
        const uint32_t w = dropout_word((uint64_t)(off >> 3) * 4 + j, key);
        f.x = (w & 0xffffu) >= thr16 ? f.x * scale : 0.f;
        f.y = (w >> 16) >= thr16 ? f.y * scale : 0.f;
      }
      o[j] = pack_bf16(f.x + r.x, f.y + r.y);  // the bf16 value the LayerNorm input holds in memory
      const float2 q = unpack_bf16(o[j]);
      v[k][2 * j] = q.x;
      v[k][2 * j + 1] = q.y;
      s += q.x + q.y;
    }
    if (sum_out != nullptr) *reinterpret_cast<uint4*>(sum_out + off) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  const float mu = warp_sum(s) * (1.f / N);
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[k][j] - mu;
      sq += d * d;
    }
  const float rs = rsqrtf(warp_sum(sq) * (1.f / N) + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const long long off = row * N + 256 * k + 8 * lane;
    uint32_t u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      u[j] = pack_bf16((v[k][2 * j] - mu) * rs * ga[k][2 * j] + be[k][2 * j],
                       (v[k][2 * j + 1] - mu) * rs * ga[k][2 * j + 1] + be[k][2 * j + 1]);
    *reinterpret_cast<uint4*>(y + off) = make_uint4(u[0], u[1], u[2], u[3]);
    if (y_add != nullptr) {
      const uint4 ua = *reinterpret_cast<const uint4*>(add + off);
      const uint32_t* pa = &ua.x;
      uint32_t t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16(u[j]), b = unpack_bf16(pa[j]);
        t[j] = pack_bf16(a.x + b.x, a.y + b.y);
      }
      *reinterpret_cast<uint4*>(y_add + off) = make_uint4(t[0], t[1], t[2], t[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ L2 normalise
// y = x / max(||x||, eps) per row (fp32); bwd: dx = (dy - y * (y . dy)) / max(||x||, eps)
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ nrm,
                                  long long rows, int N, float eps) {
  pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < N; c += 32) s += x[row * N + c] * x[row * N + c];
  const float n = fmaxf(sqrtf(warp_sum(s)), eps);
  if (lane == 0) nrm[row] = n;
  for (int c = lane; c < N; c += 32) y[row * N + c] = x[row * N + c] / n;
}
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                  const float* __restrict__ nrm, float* __restrict__ dx, long long rows, int N) {
  pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < N; c += 32) s += dy[row * N + c] * y[row * N + c];
  s = warp_sum(s);
  const float inv = 1.f / nrm[row];
  for (int c = lane; c < N; c += 32) dx[row * N + c] = (dy[row * N + c] - y[row * N + c] * s) * inv;
}

// ------------------------------------------------------------------------------------------------ attention softmax
// scores fp32 [rows, ld_s] (already scaled) -> P bf16 [rows, ld_p]; rows = B*H*Sq, key mask [B, Sk] (1 = padded).
// With dropout (pd != null): P keeps the un-dropped probabilities (needed by the backward pass), pd = dropout(P) is
// what the PV product consumes (nn.MultiheadAttention applies dropout to the attention weights).
__global__ void attn_softmax_fwd_kernel(const float* __restrict__ s, const uint8_t* __restrict__ kmask,
                                        __nv_bfloat16* __restrict__ p, __nv_bfloat16* __restrict__ pd, long long rows,
                                        int Sk, int ld_s, int ld_p, int rows_per_batch,
                                        const unsigned long long* __restrict__ seed, uint32_t site, uint32_t thr,
                                        float keep_scale) {
  pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + row * ld_s;
  const uint8_t* km = kmask ? kmask + (row / rows_per_batch) * Sk : nullptr;
  float mx = -INFINITY;
  for (int c = lane; c < Sk; c += 32) {
    const float v = (km && km[c]) ? -INFINITY : sr[c];
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < Sk; c += 32) {
    const float v = (km && km[c]) ? -INFINITY : sr[c];
    sum += expf(v - mx);  // all keys masked -> NaN, as in the reference
  }
  sum = warp_sum(sum);
  __nv_bfloat16* pr = p + row * ld_p;
  __nv_bfloat16* pdr = pd ? pd + row * ld_p : nullptr;
  const uint64_t key = pd ? dropout_key(seed, site) : 0ull;
  for (int c = lane; c < ld_p; c += 32) {
    float o = 0.f;
    if (c < Sk) {
      const float v = (km && km[c]) ? -INFINITY : sr[c];
      o = expf(v - mx) / sum;
    }
    pr[c] = __float2bfloat16_rn(o);
    if (pdr) pdr[c] = __float2bfloat16_rn(dropout_keep((uint64_t)(row * ld_p + c), key, thr) ? o * keep_scale : 0.f);
  }
}

// dS = P * (dP - sum_k dP * P) * scale   (dP fp32 [rows, ld_s], P bf16, dS bf16 [rows, ld_p]); with dropout the incoming
// gradient is w.r.t. dropout(P): dP = keep ? dPd / (1 - p) : 0 (mask regenerated from the seed)
__global__ void attn_softmax_bwd_kernel(const float* __restrict__ dp, const __nv_bfloat16* __restrict__ p,
                                        __nv_bfloat16* __restrict__ ds, long long rows, int Sk, int ld_s, int ld_p,
                                        float scale, const unsigned long long* __restrict__ seed, uint32_t site,
                                        uint32_t thr, float keep_scale) {
  pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* dr = dp + row * ld_s;
  const __nv_bfloat16* pr = p + row * ld_p;
  const uint64_t key = seed ? dropout_key(seed, site) : 0ull;
  float dot = 0.f;
  for (int c = lane; c < Sk; c += 32) {
    float g = dr[c];
    if (seed) g = dropout_keep((uint64_t)(row * ld_p + c), key, thr) ? g * keep_scale : 0.f;
    dot += g * __bfloat162float(pr[c]);
  }
  dot = warp_sum(dot);
  __nv_bfloat16* o = ds + row * ld_p;
  for (int c = lane; c < ld_p; c += 32) {
    float v = 0.f;
    if (c < Sk) {
      float g = dr[c];
      if (seed) g = dropout_keep((uint64_t)(row * ld_p + c), key, thr) ? g * keep_scale : 0.f;
      v = __bfloat162float(pr[c]) * (g - dot) * scale;
    }
    o[c] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------ sine position
// mask [B,H,W] (1 = padded) -> pos fp32 + bf16 in sequence layout [(y*W+x), B, 2F]; first F = y, last F = x
__global__ void pos_sine_kernel(const uint8_t* __restrict__ mask, float* __restrict__ pos32,
                                __nv_bfloat16* __restrict__ pos16, int B, int H, int W, int F, float temperature,
                                long long ld_rows /* B*2F */) {
  pdl_prologue();
  extern __shared__ float emb[];  // y_embed[H*W], x_embed[H*W]
  float* ye = emb;
  float* xe = emb + H * W;
  const int b = blockIdx.x;
  const uint8_t* m = mask + (size_t)b * H * W;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    float run = 0.f;
    for (int y = 0; y < H; ++y) {
      run += m[y * W + x] ? 0.f : 1.f;
      ye[y * W + x] = run;
    }
  }
  for (int y = threadIdx.x; y < H; y += blockDim.x) {
    float run = 0.f;
    for (int x = 0; x < W; ++x) {
      run += m[y * W + x] ? 0.f : 1.f;
      xe[y * W + x] = run;
    }
  }
  __syncthreads();
  const float scale = 6.283185307179586f, eps = 1e-6f;
  // blockIdx.y slices the output; every slice recomputes the (cheap) cumulative sums above
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < H * W * 2 * F; i += gridDim.y * blockDim.x) {
    const int c = i % (2 * F);
    const int pix = i / (2 * F);
    const int y = pix / W, x = pix % W;
    const bool is_y = c < F;
    const int k = is_y ? c : c - F;
    const float e = is_y ? ye[pix] / (ye[(H - 1) * W + x] + eps) * scale : xe[pix] / (xe[y * W + (W - 1)] + eps) * scale;
    const float dim_t = powf(temperature, (float)(2 * (k / 2)) / (float)F);
    const float v = e / dim_t;
    const float o = (k & 1) ? cosf(v) : sinf(v);
    const long long off = (long long)pix * ld_rows + (long long)b * 2 * F + c;
    if (pos32) pos32[off] = o;
    if (pos16) pos16[off] = __float2bfloat16_rn(o);
  }
}

// ------------------------------------------------------------------------------------------------ RoBERTa embeddings
// x[b*L + l, :] = word[id] + pos[pos_id] + type[0];  pos_id = cumsum(id != pad)[l] * (id != pad) + pad
__global__ void embed_gather_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                    const float* __restrict__ posw, const float* __restrict__ type0,
                                    float* __restrict__ out, int* __restrict__ pos_ids, int B, int L, int E, int pad_id,
                                    int seq_first) {
  pdl_prologue();
  const int b = blockIdx.x / L, l = blockIdx.x % L;
  const long long id = ids[blockIdx.x];
  const int row = seq_first ? l * B + b : blockIdx.x;  // output row (ids are always [B, L])
  int cnt = 0;
  for (int k = 0; k <= l; ++k) cnt += ids[(long long)b * L + k] != pad_id;
  const int pid = (id != pad_id ? cnt : 0) + pad_id;
  if (threadIdx.x == 0 && pos_ids) pos_ids[row] = pid;
  for (int c = threadIdx.x; c < E; c += blockDim.x)
    out[(long long)row * E + c] = word[id * E + c] + posw[(long long)pid * E + c] + type0[c];
}

template <typename T>
__global__ void embed_scatter_kernel(const T* __restrict__ dx, const long long* __restrict__ ids,
                                     const int* __restrict__ pos_ids, float* __restrict__ dword,
                                     float* __restrict__ dpos, float* __restrict__ dtype0, int B, int L, int E,
                                     int seq_first, int pad_id) {
  pdl_prologue();
  const int b = blockIdx.x / L, l = blockIdx.x % L;
  const int row = seq_first ? l * B + b : blockIdx.x;
  const long long id = ids[blockIdx.x];
  const int pid = pos_ids[row];
  // nn.Embedding(padding_idx = pad_id): the <pad> rows of the word and position tables receive no gradient
  const bool w_ok = dword != nullptr && id != (long long)pad_id;
  const bool p_ok = dpos != nullptr && pid != pad_id;
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    const float g = ld_as_float(dx, (long long)row * E + c);
    if (w_ok) atomicAdd(dword + id * E + c, g);
    if (p_ok) atomicAdd(dpos + (long long)pid * E + c, g);
    if (dtype0) atomicAdd(dtype0 + c, g);
  }
}

// Data-parallel exchange of the word-embedding gradient as (id, row) pairs instead of the dense [vocab, E] table
// (154 MB of which at most B * L rows are non-zero): `ids` / `rows` hold the pairs of ALL ranks, gathered in rank
// order.  table[id] = scale * sum of the rows with that id, summed in list order by the block of the id's FIRST
// occurrence and WRITTEN (not added): deterministic, so every rank ends up with bit-identical gradients, as after an
// all-reduce.  Rows of ids nobody holds stay as they are (zero).  grid = n_rows, block = 256.
__global__ void embed_rows_merge_kernel(float* __restrict__ table, const long long* __restrict__ ids,
                                        const float* __restrict__ rows, int n_rows, int E, long long pad_id,
                                        float scale) {
  pdl_prologue();
  __shared__ int s_list[1024];
  __shared__ int s_n;
  const int j = blockIdx.x;
  const long long id = ids[j];
  if (id == pad_id) return;
  if (threadIdx.x == 0) {
    int n = 0;
    bool first = true;
    for (int i = 0; i < j; ++i)
      if (ids[i] == id) {
        first = false;
        break;
      }
    if (first)
      for (int i = j; i < n_rows && n < 1024; ++i)
        if (ids[i] == id) s_list[n++] = i;
    s_n = first ? n : 0;
  }
  __syncthreads();
  const int n = s_n;
  if (n == 0) return;
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < n; ++k) acc += rows[(long long)s_list[k] * E + c];
    table[id * E + c] = acc * scale;
  }
}

}  // namespace toist

using namespace toist;

extern "C" {

int toist_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta, void* y_bf16,
                        float* y_f32, float* mean, float* rstd, int64_t rows, int32_t n, float eps, void* stream) {
  TOIST_REQUIRE(x && gamma && beta && (y_bf16 || y_f32), "toist_layernorm_fwd: null pointer");
  TOIST_REQUIRE(n >= 1 && n <= 32 * kMaxPerLane, "toist_layernorm_fwd: width %d unsupported (max 1024)", n);
  if (rows == 0) return TOIST_OK;
  if (x_dtype == TOIST_BF16 && y_bf16 != nullptr && y_f32 == nullptr && n % 256 == 0 && n <= 768 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y_bf16) | reinterpret_cast<uintptr_t>(gamma) |
        reinterpret_cast<uintptr_t>(beta)) & 15) == 0) {
    const unsigned g4 = (unsigned)((rows + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
#define LN_FWD_VEC(KK)                                                                                               \
  launch_pdl((layernorm_fwd_vec_kernel<KK>), dim3(g4), dim3(128), 0, st, (const __nv_bfloat16*)x, gamma, beta,       \
             (__nv_bfloat16*)y_bf16, mean, rstd, rows, eps)
    if (n == 256) LN_FWD_VEC(1);
    else if (n == 512) LN_FWD_VEC(2);
    else LN_FWD_VEC(3);
#undef LN_FWD_VEC
    TOIST_CHECK_CUDA(cudaGetLastError());
    return TOIST_OK;
  }
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (x_dtype == TOIST_F32)
    launch_pdl((layernorm_fwd_kernel<float>), dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)x, gamma, beta,
                                                                        (__nv_bfloat16*)y_bf16, y_f32, mean, rstd, rows, n, eps);
  else
    launch_pdl((layernorm_fwd_kernel<__nv_bfloat16>), dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)y_bf16, y_f32, mean, rstd, rows, n, eps);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

// dy (and optional dy2, same dtype) are summed; dx dtype selectable; dgamma/dbeta accumulate (may be null together)
static int ln_bwd_vec_launch(const void* dy, const void* dy2, const void* x, const float* mean, const float* rstd,
                             const float* gamma, void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t n,
                             void* dx_drop, const uint64_t* seed, uint32_t site, float p_drop, cudaStream_t st) {
  unsigned g4 = (unsigned)((rows + 3) / 4);
  if (g4 > 592) g4 = 592;  // 4 CTAs of 4 warps per SM; each warp then walks rows with its column sums in registers
  const uint32_t thr16 = (uint32_t)((double)p_drop * 4294967296.0) >> 16;
  const float scale = 1.f / (1.f - p_drop);
#define LN_BWD_VEC(KK)                                                                                               \
  launch_pdl((layernorm_bwd_vec_kernel<KK>), dim3(g4), dim3(128), 0, st, (const __nv_bfloat16*)dy,                   \
             (const __nv_bfloat16*)dy2, (const __nv_bfloat16*)x, mean, rstd, gamma, (__nv_bfloat16*)dx, dgamma,      \
             dbeta, (long long)rows, (__nv_bfloat16*)dx_drop, (const unsigned long long*)seed, site, thr16, scale)
  if (n == 256) LN_BWD_VEC(1);
  else if (n == 512) LN_BWD_VEC(2);
  else LN_BWD_VEC(3);
#undef LN_BWD_VEC
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

static bool ln_vec_ok(int32_t n, std::initializer_list<const void*> ptrs) {
  if (n % 256 != 0 || n > 768) return false;
  uintptr_t al = 0;
  for (const void* q : ptrs) al |= reinterpret_cast<uintptr_t>(q);
  return (al & 15) == 0;
}

// s = res + dropout(x) (p_drop = 0 / seed null: no dropout; res null: none), y = LayerNorm(s), y_add = y + add;
// sum_out (nullable) receives s.  bf16 rows of 256 / 512 / 768 elements, 16-byte aligned (TOIST_ERR_UNSUPPORTED otherwise:
// the caller then issues the separate kernels).
int toist_layernorm_fused_fwd(const void* x, const void* res, const void* add, const float* gamma, const float* beta,
                              void* sum_out, void* y, void* y_add, float* mean, float* rstd, int64_t rows, int32_t n,
                              float eps, float p_drop, const uint64_t* seed, uint32_t site, void* stream) {
  TOIST_REQUIRE(x && gamma && beta && y, "toist_layernorm_fused_fwd: null pointer");
  TOIST_REQUIRE((add == nullptr) == (y_add == nullptr), "toist_layernorm_fused_fwd: add needs y_add");
  TOIST_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || seed != nullptr),
                "toist_layernorm_fused_fwd: dropout needs a seed and 0 <= p < 1");
  if (!ln_vec_ok(n, {x, res, add, gamma, beta, sum_out, y, y_add}))
    return set_error(TOIST_ERR_UNSUPPORTED, "toist_layernorm_fused_fwd: width %d / alignment not supported", n);
  if (rows == 0) return TOIST_OK;
  const uint32_t thr16 = (uint32_t)((double)p_drop * 4294967296.0) >> 16;
  const float scale = 1.f / (1.f - p_drop);
  const unsigned long long* sd = p_drop > 0.f ? (const unsigned long long*)seed : nullptr;
  const unsigned g4 = (unsigned)((rows + 3) / 4);
  cudaStream_t st = (cudaStream_t)stream;
#define LN_FUSED(KK)                                                                                                  \
  launch_pdl((layernorm_fused_fwd_vec_kernel<KK>), dim3(g4), dim3(128), 0, st, (const __nv_bfloat16*)x,              \
             (const __nv_bfloat16*)res, (const __nv_bfloat16*)add, gamma, beta, (__nv_bfloat16*)sum_out,              \
             (__nv_bfloat16*)y, (__nv_bfloat16*)y_add, mean, rstd, (long long)rows, eps, sd, site, thr16, scale)
  if (n == 256) LN_FUSED(1);
  else if (n == 512) LN_FUSED(2);
  else LN_FUSED(3);
#undef LN_FUSED
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

// LayerNorm backward (as toist_layernorm_bwd, bf16) that also writes dx_drop = dropout-backward(dx) for site `site`
int toist_layernorm_bwd_drop(const void* dy, const void* dy2, const void* x, const float* mean, const float* rstd,
                             const float* gamma, void* dx, void* dx_drop, float* dgamma, float* dbeta, int64_t rows,
                             int32_t n, float p_drop, const uint64_t* seed, uint32_t site, void* stream) {
  TOIST_REQUIRE(dy && x && mean && rstd && gamma && dx && dx_drop && seed, "toist_layernorm_bwd_drop: null pointer");
  TOIST_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "toist_layernorm_bwd_drop: pass both dgamma and dbeta or neither");
  TOIST_REQUIRE(p_drop > 0.f && p_drop < 1.f, "toist_layernorm_bwd_drop: 0 < p < 1");
  if (!ln_vec_ok(n, {dy, dy2, x, gamma, dx, dx_drop}))
    return set_error(TOIST_ERR_UNSUPPORTED, "toist_layernorm_bwd_drop: width %d / alignment not supported", n);
  if (rows == 0) return TOIST_OK;
  return ln_bwd_vec_launch(dy, dy2, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, n, dx_drop, seed, site, p_drop,
                           (cudaStream_t)stream);
}

int toist_layernorm_bwd(const void* dy, const void* dy2, int32_t dy_dtype, const void* x, int32_t x_dtype,
                        const float* mean, const float* rstd, const float* gamma, void* dx, int32_t dx_dtype,
                        float* dgamma, float* dbeta, int64_t rows, int32_t n, void* stream) {
  TOIST_REQUIRE(dy && x && mean && rstd && gamma && dx, "toist_layernorm_bwd: null pointer");
  TOIST_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "toist_layernorm_bwd: pass both dgamma and dbeta or neither");
  TOIST_REQUIRE(n >= 1 && n <= 32 * kMaxPerLane, "toist_layernorm_bwd: width %d unsupported (max 1024)", n);
  if (rows == 0) return TOIST_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == TOIST_BF16 && dy_dtype == TOIST_BF16 && dx_dtype == TOIST_BF16 && ln_vec_ok(n, {x, dy, dy2, dx, gamma}))
    return ln_bwd_vec_launch(dy, dy2, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, n, nullptr, nullptr, 0u, 0.f, st);
  unsigned grid = (unsigned)((rows + 7) / 8);
  if (grid > 296) grid = 296;
  const size_t smem = 2 * (size_t)n * sizeof(float);
#define LN_BWD(TX, TDY, TDX)                                                                                      \
  launch_pdl((layernorm_bwd_kernel<TX, TDY, TDX>), dim3(grid), dim3(256), smem, st, (const TDY*)dy, (const TDY*)dy2, (const TX*)x, mean, \
                                                              rstd, gamma, (TDX*)dx, dgamma, dbeta, rows, n)
  const int key = (x_dtype << 2) | (dy_dtype << 1) | dx_dtype;
  switch (key) {
    case 0: LN_BWD(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16); break;
    case 1: LN_BWD(__nv_bfloat16, __nv_bfloat16, float); break;
    case 2: LN_BWD(__nv_bfloat16, float, __nv_bfloat16); break;
    case 3: LN_BWD(__nv_bfloat16, float, float); break;
    case 4: LN_BWD(float, __nv_bfloat16, __nv_bfloat16); break;
    case 5: LN_BWD(float, __nv_bfloat16, float); break;
    case 6: LN_BWD(float, float, __nv_bfloat16); break;
    default: LN_BWD(float, float, float); break;
  }
#undef LN_BWD
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_l2norm_fwd(const float* x, float* y, float* nrm, int64_t rows, int32_t n, float eps, void* stream) {
  TOIST_REQUIRE(x && y && nrm, "toist_l2norm_fwd: null pointer");
  if (rows == 0) return TOIST_OK;
  launch_pdl(l2norm_fwd_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, x, y, nrm, rows, n, eps);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_l2norm_bwd(const float* dy, const float* y, const float* nrm, float* dx, int64_t rows, int32_t n,
                     void* stream) {
  TOIST_REQUIRE(dy && y && nrm && dx, "toist_l2norm_bwd: null pointer");
  if (rows == 0) return TOIST_OK;
  launch_pdl(l2norm_bwd_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, dy, y, nrm, dx, rows, n);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_attn_softmax_fwd(const float* scores, const uint8_t* key_mask, void* probs, void* probs_dropped,
                           int64_t rows, int32_t sk, int32_t ld_s, int32_t ld_p, int32_t rows_per_batch, float p_drop,
                           const uint64_t* seed, uint32_t site, void* stream) {
  TOIST_REQUIRE(scores && probs && rows_per_batch > 0, "toist_attn_softmax_fwd: bad arguments");
  TOIST_REQUIRE(probs_dropped == nullptr || (seed != nullptr && p_drop > 0.f && p_drop < 1.f),
                "toist_attn_softmax_fwd: dropout needs a seed and 0 < p < 1");
  if (rows == 0) return TOIST_OK;
  const uint32_t thr = (uint32_t)((double)p_drop * 4294967296.0);
  launch_pdl(attn_softmax_fwd_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, 
      scores, key_mask, (__nv_bfloat16*)probs, (__nv_bfloat16*)probs_dropped, rows, sk, ld_s, ld_p, rows_per_batch,
      (const unsigned long long*)seed, site, thr, 1.f / (1.f - p_drop));
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_attn_softmax_bwd(const float* dprobs, const void* probs, void* dscores, int64_t rows, int32_t sk,
                           int32_t ld_s, int32_t ld_p, float scale, float p_drop, const uint64_t* seed, uint32_t site,
                           void* stream) {
  TOIST_REQUIRE(dprobs && probs && dscores, "toist_attn_softmax_bwd: null pointer");
  if (rows == 0) return TOIST_OK;
  const bool drop = seed != nullptr && p_drop > 0.f;
  const uint32_t thr = (uint32_t)((double)p_drop * 4294967296.0);
  launch_pdl(attn_softmax_bwd_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, 
      dprobs, (const __nv_bfloat16*)probs, (__nv_bfloat16*)dscores, rows, sk, ld_s, ld_p, scale,
      drop ? (const unsigned long long*)seed : nullptr, site, thr, 1.f / (1.f - p_drop));
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_pos_sine(const uint8_t* mask, float* pos_f32, void* pos_bf16, int32_t batch, int32_t h, int32_t w,
                   int32_t num_pos_feats, float temperature, void* stream) {
  TOIST_REQUIRE(mask && (pos_f32 || pos_bf16), "toist_pos_sine: null pointer");
  const size_t smem = 2 * (size_t)h * w * sizeof(float);
  TOIST_REQUIRE(smem <= 200 * 1024, "toist_pos_sine: feature map %dx%d too large", h, w);
  static bool configured = false;
  if (!configured) {
    TOIST_CHECK_CUDA(cudaFuncSetAttribute(pos_sine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  const int slices = (h * w * 2 * num_pos_feats + 4095) / 4096 < 32 ? (h * w * 2 * num_pos_feats + 4095) / 4096 : 32;
  launch_pdl(pos_sine_kernel, dim3(dim3(batch, slices > 0 ? slices : 1)), dim3(256), smem, (cudaStream_t)stream, mask, pos_f32, (__nv_bfloat16*)pos_bf16, batch, h, w,
                                                              num_pos_feats, temperature,
                                                              (long long)batch * 2 * num_pos_feats);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_embed_gather(const int64_t* ids, const float* word, const float* pos, const float* type0, float* out,
                       int32_t* pos_ids, int32_t batch, int32_t len, int32_t dim, int32_t pad_id, int32_t seq_first,
                       void* stream) {
  TOIST_REQUIRE(ids && word && pos && type0 && out, "toist_embed_gather: null pointer");
  if (batch * len == 0) return TOIST_OK;
  launch_pdl(embed_gather_kernel, dim3(batch * len), dim3(256), 0, (cudaStream_t)stream, (const long long*)ids, word, pos, type0, out,
                                                                     pos_ids, batch, len, dim, pad_id, seq_first);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_embed_scatter(const void* dx, int32_t dx_dtype, const int64_t* ids, const int32_t* pos_ids, float* dword,
                        float* dpos, float* dtype0, int32_t batch, int32_t len, int32_t dim, int32_t seq_first,
                        int32_t pad_id, void* stream) {
  TOIST_REQUIRE(dx && ids && pos_ids, "toist_embed_scatter: null pointer");
  const int rows = batch * len;
  if (rows == 0) return TOIST_OK;
  if (dx_dtype == TOIST_BF16)
    launch_pdl((embed_scatter_kernel<__nv_bfloat16>), dim3(rows), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)dx, (const long long*)ids, pos_ids, dword, dpos, dtype0, batch, len, dim, seq_first, pad_id);
  else
    launch_pdl((embed_scatter_kernel<float>), dim3(rows), dim3(256), 0, (cudaStream_t)stream, (const float*)dx, (const long long*)ids, pos_ids, dword, dpos, dtype0, batch, len, dim, seq_first, pad_id);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_embed_rows_merge(float* table, const int64_t* ids, const float* rows, int32_t n_rows, int32_t dim,
                           int64_t pad_id, float scale, void* stream) {
  TOIST_REQUIRE(table && ids && rows, "toist_embed_rows_merge: null pointer");
  TOIST_REQUIRE(n_rows >= 0 && n_rows <= 65535 && dim >= 1, "toist_embed_rows_merge: bad sizes");
  if (n_rows == 0) return TOIST_OK;
  launch_pdl(embed_rows_merge_kernel, dim3((unsigned)n_rows), dim3(256), 0, (cudaStream_t)stream, table,
             (const long long*)ids, rows, (int)n_rows, (int)dim, (long long)pad_id, scale);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

}  // extern "C"
