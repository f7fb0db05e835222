// HBM-bound helper kernels: casts / weight preparation, stem im2col, max-pool, bias-gradient column sums,
// activation derivatives, broadcasts.  All vectorised to 16-byte accesses where the layout allows it.
#include "common.cuh"
#include "host_util.h"

namespace toist {

// ------------------------------------------------------------------------------------------------ weight prep
// dst[r, c] = bf16(src[r, c] * row_scale[r])  for one or many [rows, cols] fp32 matrices (dst leading dim ldd).
struct PrepItem {
  const float* src;
  __nv_bfloat16* dst;
  const float* row_scale;  // may be null
  int rows, cols, ldd;
  int first_block;  // prefix of blocks (each block handles 2048 elements)
  int taps;         // > 1: src is [rows][cols / taps][taps] (conv OIHW) and dst gets [rows][taps][cols / taps] (OHWI)
  int pad_;
};

__global__ void weight_prep_kernel(const PrepItem* __restrict__ items, int n_items) {
  pdl_prologue();
  // binary search the item owning this block
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PrepItem it = items[lo];
  const long long total = (long long)it.rows * it.cols;
  const long long base = (long long)(blockIdx.x - it.first_block) * 2048;
  if (it.ldd == it.cols && it.row_scale == nullptr && (it.cols % 8 == 0) && it.taps <= 1) {
    const long long i = base + (long long)threadIdx.x * 8;
    if (i + 8 <= total) {
      const float4 a = *reinterpret_cast<const float4*>(it.src + i);
      const float4 b = *reinterpret_cast<const float4*>(it.src + i + 4);
      uint4 u;
      u.x = pack_bf16(a.x, a.y); u.y = pack_bf16(a.z, a.w); u.z = pack_bf16(b.x, b.y); u.w = pack_bf16(b.z, b.w);
      *reinterpret_cast<uint4*>(it.dst + i) = u;
      return;
    }
  }
  for (int k = 0; k < 8; ++k) {
    const long long i = base + (long long)threadIdx.x * 8 + k;
    if (i >= total) break;
    const int r = (int)(i / it.cols), c = (int)(i % it.cols);  // c indexes the destination row
    const float s = it.row_scale ? it.row_scale[r] : 1.f;
    long long si = i;
    if (it.taps > 1) {
      const int cin = it.cols / it.taps;
      const int t = c / cin, ci = c % cin;
      si = (long long)r * it.cols + (long long)ci * it.taps + t;
    }
    it.dst[(long long)r * it.ldd + c] = __float2bfloat16_rn(it.src[si] * s);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_prologue();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i);
    const float4 b = *reinterpret_cast<const float4*>(src + i + 4);
    uint4 u;
    u.x = pack_bf16(a.x, a.y); u.y = pack_bf16(a.z, a.w); u.z = pack_bf16(b.x, b.y); u.w = pack_bf16(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + i) = u;
  } else {
    for (long long k = i; k < n; ++k) dst[k] = __float2bfloat16_rn(src[k]);
  }
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  pdl_prologue();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + i);
    const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
    *reinterpret_cast<float4*>(dst + i) = make_float4(f0.x, f0.y, f1.x, f1.y);
    *reinterpret_cast<float4*>(dst + i + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
  } else {
    for (long long k = i; k < n; ++k) dst[k] = __bfloat162float(src[k]);
  }
}

// out = a + b (bf16), optional third addend c; 8 elements per thread
__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                const __nv_bfloat16* __restrict__ c, __nv_bfloat16* __restrict__ out, long long n) {
  pdl_prologue();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const uint4 ua = *reinterpret_cast<const uint4*>(a + i), ub = *reinterpret_cast<const uint4*>(b + i);
    uint4 uc = make_uint4(0, 0, 0, 0);
    if (c) uc = *reinterpret_cast<const uint4*>(c + i);
    const uint32_t* pa = &ua.x; const uint32_t* pb = &ub.x; const uint32_t* pc = &uc.x;
    uint4 uo; uint32_t* po = &uo.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack_bf16(pa[k]), fb = unpack_bf16(pb[k]), fc = unpack_bf16(pc[k]);
      po[k] = pack_bf16(fa.x + fb.x + fc.x, fa.y + fb.y + fc.y);
    }
    *reinterpret_cast<uint4*>(out + i) = uo;
  } else {
    for (long long k = i; k < n; ++k)
      out[k] = __float2bfloat16_rn(__bfloat162float(a[k]) + __bfloat162float(b[k]) + (c ? __bfloat162float(c[k]) : 0.f));
  }
}

// ------------------------------------------------------------------------------------------------ stem im2col
// images fp32 NCHW [N,3,H,W] -> patches bf16 [N*Ho*Wo, ldk] for the 7x7/2 pad-3 stem conv; column = (ky*7+kx)*3+c.
// One CTA per strip of 64 output pixels of one output row: the 7 input rows x 133 input columns x 3 channels the strip
// touches are staged in shared memory with coalesced fp32 reads (converted to bf16 once), then every thread assembles
// 16-byte chunks (8 patch columns) of the output rows from a column -> shared-offset table and stores them coalesced.
constexpr int kStemStrip = 64;
constexpr int kStemCols = 2 * kStemStrip + 5;  // input columns touched by a strip (stride 2, 7 taps)
// Stem input for the implicit-GEMM path: fp32 NCHW [n, 3, h, w] -> bf16 [n, hp, wp, 8] (channels 3..7 zero) with a zero
// border of 3 pixels on top / left (and at least 3 on the bottom / right): one 16-byte store per output pixel.  A 7 x 7
// stride-2 window row of output pixel x is then the 64 contiguous elements starting at padded pixel 2 x (7 pixels x 8
// channels + one pixel that meets zero weights), which a TMA box with overlapping rows (row pitch 2 pixels) fetches
// directly: the 315 MB im2col matrix of the round-1 stem (written once, read once) is never materialised.
__global__ void stem_pad_nhwc8_kernel(const float* __restrict__ img, uint4* __restrict__ out, int N, int H, int W, int Hp,
                                      int Wp) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * Hp * Wp;
  if (i >= total) return;
  const int xp = (int)(i % Wp);
  const long long t = i / Wp;
  const int yp = (int)(t % Hp);
  const int n = (int)(t / Hp);
  const int x = xp - 3, y = yp - 3;
  uint4 u = make_uint4(0u, 0u, 0u, 0u);
  if (x >= 0 && x < W && y >= 0 && y < H) {
    const long long plane = (long long)H * W;
    const float* p = img + (long long)n * 3 * plane + (long long)y * W + x;
    u.x = pack_bf16(__ldg(p), __ldg(p + plane));
    u.y = pack_bf16(__ldg(p + 2 * plane), 0.f);
  }
  out[i] = u;
}

__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int N, int H, int W, int Ho, int Wo,
                   int ldk) {
  __shared__ __nv_bfloat16 tile[3 * 7 * (kStemCols + 3)];
  __shared__ short lut[256];
  constexpr int kPitch = kStemCols + 3;
  const int strips = (Wo + kStemStrip - 1) / kStemStrip;
  int t = blockIdx.x;
  const int sx = t % strips;
  t /= strips;
  const int oy = t % Ho;
  const int n = t / Ho;
  const int ox0 = sx * kStemStrip;
  for (int col = threadIdx.x; col < ldk && col < 256; col += blockDim.x) {
    short off = -1;
    if (col < 147) {
      const int c = col % 3, kx = (col / 3) % 7, ky = col / 21;
      off = (short)((c * 7 + ky) * kPitch + kx);
    }
    lut[col] = off;
  }
  pdl_prologue();
  const int ix0 = ox0 * 2 - 3, iy0 = oy * 2 - 3;
  for (int i = threadIdx.x; i < 21 * kStemCols; i += blockDim.x) {
    const int r = i / kStemCols, dx = i - r * kStemCols;  // r = c * 7 + ky
    const int c = r / 7, ky = r - c * 7;
    const int iy = iy0 + ky, ix = ix0 + dx;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + (((long long)n * 3 + c) * H + iy) * W + ix);
    tile[r * kPitch + dx] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int chunks = ldk >> 3;
  const int npx = min(kStemStrip, Wo - ox0);
  __nv_bfloat16* obase = out + (((long long)n * Ho + oy) * Wo + ox0) * ldk;
  const unsigned short* tl = reinterpret_cast<const unsigned short*>(tile);
  for (int i = threadIdx.x; i < npx * chunks; i += blockDim.x) {
    const int px = i / chunks, ch = i - px * chunks;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const short o0 = lut[ch * 8 + 2 * j], o1 = lut[ch * 8 + 2 * j + 1];
      const uint32_t lo = o0 >= 0 ? tl[o0 + 2 * px] : 0u;
      const uint32_t hi = o1 >= 0 ? tl[o1 + 2 * px] : 0u;
      w[j] = lo | (hi << 16);
    }
    *reinterpret_cast<uint4*>(obase + (long long)px * ldk + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ------------------------------------------------------------------------------------------------ max pool 3x3/2 pad 1
// NHWC bf16; each thread handles 8 channels of one output pixel
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                    int W, int C, int Ho, int Wo) {
  pdl_prologue();
  const int cg = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * Ho * Wo * cg) return;
  const int c8 = (int)(idx % cg);
  long long p = idx / cg;
  const int ox = (int)(p % Wo); p /= Wo;
  const int oy = (int)(p % Ho);
  const int n = (int)(p / Ho);
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = oy * 2 + dy - 1;
    if (iy < 0 || iy >= H) continue;
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = ox * 2 + dx - 1;
      if (ix < 0 || ix >= W) continue;
      const uint4 u = *reinterpret_cast<const uint4*>(x + (((long long)n * H + iy) * W + ix) * C + c8 * 8);
      const uint32_t* pu = &u.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16(pu[k]);
        m[2 * k] = fmaxf(m[2 * k], f.x);
        m[2 * k + 1] = fmaxf(m[2 * k + 1], f.y);
      }
    }
  }
  uint4 o;
  o.x = pack_bf16(m[0], m[1]); o.y = pack_bf16(m[2], m[3]); o.z = pack_bf16(m[4], m[5]); o.w = pack_bf16(m[6], m[7]);
  *reinterpret_cast<uint4*>(y + (((long long)n * Ho + oy) * Wo + ox) * C + c8 * 8) = o;
}

// ------------------------------------------------------------------------------------------------ column sums
// out[c] += sum_r x[r, c]   (bias gradients).  grid = (ceil(C/64), row_chunks), block = (64, 4)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long R, int C, long long ld,
                              int rows_per_block) {
  pdl_prologue();
  const int c = blockIdx.x * 64 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  float acc = 0.f;
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      if constexpr (sizeof(T) == 2) acc += __bfloat162float(x[r * ld + c]);
      else acc += x[r * ld + c];
    }
  __shared__ float red[4][64];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) atomicAdd(out + c, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// Vectorised variant for bf16 with 16-byte aligned rows: a warp covers 256 consecutive columns (8 per lane, one 16-byte
// load each), the 8 warps of a block take interleaved rows, partial sums meet in shared memory.
__global__ void colsum_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long R, int C,
                                       long long ld, int rows_per_block) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (c0 < C) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + r * ld + c0));
      const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
      acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
      acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
    }
  }
  __shared__ float red[8][256 + 8];
#pragma unroll
  for (int k = 0; k < 8; ++k) red[warp][lane * 8 + k] = acc[k];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// ------------------------------------------------------------------------------------------------ activation grads
// dpre = dy * gelu'(pre)   (erf GELU), bf16
__global__ void gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ pre,
                                __nv_bfloat16* __restrict__ dx, long long n) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = __bfloat162float(pre[i]);
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  dx[i] = __float2bfloat16_rn(__bfloat162float(dy[i]) * (cdf + x * pdf));
}

// dz = (dy + dy2) * (y > 0), bf16, 8 elements per thread
__global__ void relu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                                const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, long long n) {
  pdl_prologue();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const uint4 ua = *reinterpret_cast<const uint4*>(dy + i), uy = *reinterpret_cast<const uint4*>(y + i);
    uint4 ub = make_uint4(0, 0, 0, 0);
    if (dy2) ub = *reinterpret_cast<const uint4*>(dy2 + i);
    const uint32_t* pa = &ua.x; const uint32_t* pb = &ub.x; const uint32_t* py = &uy.x;
    uint4 uo; uint32_t* po = &uo.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack_bf16(pa[k]), fb = unpack_bf16(pb[k]), fy = unpack_bf16(py[k]);
      po[k] = pack_bf16(fy.x > 0.f ? fa.x + fb.x : 0.f, fy.y > 0.f ? fa.y + fb.y : 0.f);
    }
    *reinterpret_cast<uint4*>(out + i) = uo;
  } else {
    for (long long k = i; k < n; ++k) {
      const float g = __bfloat162float(dy[k]) + (dy2 ? __bfloat162float(dy2[k]) : 0.f);
      out[k] = __float2bfloat16_rn(__bfloat162float(y[k]) > 0.f ? g : 0.f);
    }
  }
}

// dx = dy * y * (1 - y), fp32 (sigmoid output y)
__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                   long long n) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dx[i] = dy[i] * y[i] * (1.f - y[i]);
}

// out[a, c] (+)= sum_r x[a, r, c]   (e.g. query_embed gradient summed over the batch); bf16 or fp32 in, fp32 out
template <typename T>
__global__ void sum_mid_kernel(const T* __restrict__ x, float* __restrict__ out, int A, int R, int C, int accumulate) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)A * C) return;
  const int a = (int)(i / C), c = (int)(i % C);
  float acc = 0.f;
  for (int r = 0; r < R; ++r) {
    if constexpr (sizeof(T) == 2) acc += __bfloat162float(x[((long long)a * R + r) * C + c]);
    else acc += x[((long long)a * R + r) * C + c];
  }
  if (accumulate) out[i] += acc; else out[i] = acc;
}

// out[a, r, c] = x[a, c]  (bf16 out; x fp32) — query_embed broadcast over the batch
__global__ void bcast_mid_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int A, int R, int C) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)A * R * C) return;
  const int c = (int)(i % C);
  const int a = (int)(i / ((long long)R * C));
  out[i] = __float2bfloat16_rn(x[(long long)a * C + c]);
}

// NCHW fp32 -> NHWC bf16 and back (boundary conversions for tensors handed to / taken from the caller)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int HW) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * C * HW) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  const int hw = (int)(p % HW);
  const int n = (int)(p / HW);
  y[i] = __float2bfloat16_rn(x[((long long)n * C + c) * HW + hw]);
}
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int N, int C, int HW) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * C * HW) return;
  const int hw = (int)(i % HW);
  const long long p = i / HW;
  const int c = (int)(p % C);
  const int n = (int)(p / C);
  y[i] = __bfloat162float(x[((long long)n * HW + hw) * C + c]);
}

// dst[r, 0:ld] = bf16(src[r, 0:n]) zero-padded to ld columns (TMA needs 16-byte rows: e.g. the 4-wide box head)
__global__ void cast_pad_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows, int n,
                                int ld) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const long long r = i / ld;
  const int c = (int)(i % ld);
  dst[i] = __float2bfloat16_rn(c < n ? src[r * n + c] : 0.f);
}

// out[i] = keep(i) ? x[i] / (1 - p) : 0  (+ res[i]);  nn.Dropout forward, and (with x = upstream gradient, res = null)
// its backward.  In place (out == x) is allowed.
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ out, long long n,
                               const unsigned long long* __restrict__ seed, uint32_t site, uint32_t thr, float scale) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = dropout_key(seed, site);
  float v;
  if constexpr (sizeof(T) == 2) v = __bfloat162float(x[i]); else v = x[i];
  v = dropout_keep((uint64_t)i, key, thr) ? v * scale : 0.f;
  if (res != nullptr) {
    if constexpr (sizeof(T) == 2) v += __bfloat162float(res[i]); else v += res[i];
  }
  if constexpr (sizeof(T) == 2) out[i] = __float2bfloat16_rn(v); else out[i] = v;
}

// bf16, 16-byte aligned, n % 8 == 0: 8 elements and 4 hashes per thread, 16-byte loads and stores
__global__ void dropout_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                        __nv_bfloat16* __restrict__ out, long long n8,
                                        const unsigned long long* __restrict__ seed, uint32_t site, uint32_t thr,
                                        float scale) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint64_t key = dropout_key(seed, site);
  const uint32_t thr16 = thr >> 16;
  const uint4 ux = *reinterpret_cast<const uint4*>(x + i * 8);
  uint4 ur = make_uint4(0u, 0u, 0u, 0u);
  if (res != nullptr) ur = *reinterpret_cast<const uint4*>(res + i * 8);
  const uint32_t* px = &ux.x;
  const uint32_t* pr = &ur.x;
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t w = dropout_word((uint64_t)i * 4 + j, key);
    const float2 f = unpack_bf16(px[j]), r = unpack_bf16(pr[j]);
    const float a = ((w & 0xffffu) >= thr16 ? f.x * scale : 0.f) + r.x;
    const float b = ((w >> 16) >= thr16 ? f.y * scale : 0.f) + r.y;
    o[j] = pack_bf16(a, b);
  }
  *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// dst[a, c, b] = src[a, b, c]  (fp32) — conv weight gradients come out of the WGRAD engine as [Cout][taps][Cin]
__global__ void permute_021_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int B, int C) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)A * B * C) return;
  const int b = (int)(i % B);
  const long long p = i / B;
  const int c = (int)(p % C);
  const int a = (int)(p / C);
  dst[i] = src[((long long)a * B + b) * C + c];
}

// Nearest-neighbour resize of the padding mask (models/backbone.py:78 -> F.interpolate default mode) and assembly
// of the encoder key-padding mask [B, h*w + L] (models/transformer.py:102,134,146).
__global__ void key_mask_kernel(const uint8_t* __restrict__ pad, const long long* __restrict__ text_attn,
                                uint8_t* __restrict__ small, uint8_t* __restrict__ key, int B, int H, int W, int h,
                                int w, int L) {
  pdl_prologue();
  const int S = h * w + L;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * S) return;
  const int b = (int)(i / S), s = (int)(i % S);
  uint8_t v;
  if (s < h * w) {
    const int y = s / w, x = s % w;
    // torch nearest: src = floor(dst * (in / out)) computed in float
    int sy = (int)floorf((float)y * ((float)H / (float)h));
    int sx = (int)floorf((float)x * ((float)W / (float)w));
    sy = sy < H - 1 ? sy : H - 1;
    sx = sx < W - 1 ? sx : W - 1;
    v = pad[((long long)b * H + sy) * W + sx] ? 1 : 0;
    if (small) small[((long long)b * h + y) * w + x] = v;
  } else {
    v = text_attn ? (text_attn[(long long)b * L + (s - h * w)] != 1 ? 1 : 0) : 0;
  }
  if (key) key[i] = v;
}

}  // namespace toist

using namespace toist;

static inline unsigned nblk(long long n, int per) { return (unsigned)((n + per - 1) / per); }

extern "C" {

// items: device array of toist PrepItem-compatible records (see toist_b200.h toist_prep_item); total_blocks = sum of
// ceil(rows*cols / 2048) over items.
int toist_weight_prep(const void* items_dev, int32_t n_items, int32_t total_blocks, void* stream) {
  TOIST_REQUIRE(items_dev != nullptr && n_items > 0 && total_blocks > 0, "toist_weight_prep: bad arguments");
  launch_pdl(weight_prep_kernel, dim3(total_blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const PrepItem*>(items_dev),
                                                                     n_items);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  TOIST_REQUIRE(src && dst, "toist_cast_f32_bf16: null pointer");
  if (n == 0) return TOIST_OK;
  launch_pdl(cast_f32_bf16_kernel, dim3(nblk(n, 2048)), dim3(256), 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst, n);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream) {
  TOIST_REQUIRE(src && dst, "toist_cast_bf16_f32: null pointer");
  if (n == 0) return TOIST_OK;
  launch_pdl(cast_bf16_f32_kernel, dim3(nblk(n, 2048)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)src, dst, n);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_add_bf16(const void* a, const void* b, const void* c, void* out, int64_t n, void* stream) {
  TOIST_REQUIRE(a && b && out, "toist_add_bf16: null pointer");
  if (n == 0) return TOIST_OK;
  launch_pdl(add_bf16_kernel, dim3(nblk(n, 2048)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                   (const __nv_bfloat16*)c, (__nv_bfloat16*)out, n);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_stem_im2col(const float* images, void* patches, int32_t n, int32_t h, int32_t w, int32_t ldk, void* stream) {
  TOIST_REQUIRE(images && patches && ldk >= 147 && ldk % 8 == 0, "toist_stem_im2col: bad arguments");
  const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1;
  const long long pix = (long long)n * ho * wo;
  TOIST_REQUIRE(ldk <= 256, "toist_stem_im2col: ldk %d too wide", ldk);
  (void)pix;
  const long long blocks = (long long)n * ho * ((wo + kStemStrip - 1) / kStemStrip);
  launch_pdl(stem_im2col_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, images,
             (__nv_bfloat16*)patches, n, h, w, ho, wo, ldk);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_stem_pad_nhwc8(const float* images, void* out, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp,
                         void* stream) {
  TOIST_REQUIRE(images && out && hp >= h + 6 && wp >= w + 6, "toist_stem_pad_nhwc8: bad arguments");
  TOIST_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "toist_stem_pad_nhwc8: output must be 16-byte aligned");
  const long long total = (long long)n * hp * wp;
  launch_pdl(stem_pad_nhwc8_kernel, dim3(nblk(total, 256)), dim3(256), 0, (cudaStream_t)stream, images,
             reinterpret_cast<uint4*>(out), n, h, w, hp, wp);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_maxpool3x3s2(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, void* stream) {
  TOIST_REQUIRE(x && y && c % 8 == 0, "toist_maxpool3x3s2: channels must be a multiple of 8");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const long long total = (long long)n * ho * wo * (c / 8);
  launch_pdl(maxpool3x3s2_kernel, dim3(nblk(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y,
                                                                          n, h, w, c, ho, wo);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_colsum(const void* x, int32_t dtype, float* out, int64_t rows, int32_t cols, int64_t ld, void* stream) {
  TOIST_REQUIRE(x && out, "toist_colsum: null pointer");
  if (rows == 0) return TOIST_OK;
  if (dtype == TOIST_BF16 && cols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int col_blocks = (cols + 255) / 256;
    int chunks = (int)((rows + 63) / 64);               // at least 64 rows per block
    const int want = (592 + col_blocks - 1) / col_blocks;  // ~4 blocks per SM in total
    if (chunks > want) chunks = want;
    if (chunks < 1) chunks = 1;
    const int rpb = (int)((rows + chunks - 1) / chunks);
    dim3 grid(col_blocks, (unsigned)((rows + rpb - 1) / rpb));
    launch_pdl(colsum_bf16_vec_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, out, rows, cols, ld, rpb);
    TOIST_CHECK_CUDA(cudaGetLastError());
    return TOIST_OK;
  }
  int chunks = (int)((rows + 511) / 512);
  if (chunks > 256) chunks = 256;
  const int rpb = (int)((rows + chunks - 1) / chunks);
  dim3 grid((cols + 63) / 64, chunks), block(64, 4);
  if (dtype == TOIST_BF16)
    launch_pdl((colsum_kernel<__nv_bfloat16>), dim3(grid), dim3(block), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, out, rows, cols, ld, rpb);
  else
    launch_pdl((colsum_kernel<float>), dim3(grid), dim3(block), 0, (cudaStream_t)stream, (const float*)x, out, rows, cols, ld, rpb);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_gelu_bwd(const void* dy, const void* pre, void* dx, int64_t n, void* stream) {
  TOIST_REQUIRE(dy && pre && dx, "toist_gelu_bwd: null pointer");
  if (n == 0) return TOIST_OK;
  launch_pdl(gelu_bwd_kernel, dim3(nblk(n, 256)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)pre,
                                                                  (__nv_bfloat16*)dx, n);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_relu_bwd(const void* dy, const void* dy2, const void* y, void* out, int64_t n, void* stream) {
  TOIST_REQUIRE(dy && y && out, "toist_relu_bwd: null pointer");
  if (n == 0) return TOIST_OK;
  launch_pdl(relu_bwd_kernel, dim3(nblk(n, 2048)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2,
                                                                   (const __nv_bfloat16*)y, (__nv_bfloat16*)out, n);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  TOIST_REQUIRE(dy && y && dx, "toist_sigmoid_bwd: null pointer");
  if (n == 0) return TOIST_OK;
  launch_pdl(sigmoid_bwd_kernel, dim3(nblk(n, 256)), dim3(256), 0, (cudaStream_t)stream, dy, y, dx, n);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_sum_mid(const void* x, int32_t dtype, float* out, int32_t a, int32_t r, int32_t c, int32_t accumulate,
                  void* stream) {
  TOIST_REQUIRE(x && out, "toist_sum_mid: null pointer");
  const long long n = (long long)a * c;
  if (n == 0) return TOIST_OK;
  if (dtype == TOIST_BF16)
    launch_pdl((sum_mid_kernel<__nv_bfloat16>), dim3(nblk(n, 256)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, out, a, r, c, accumulate);
  else
    launch_pdl((sum_mid_kernel<float>), dim3(nblk(n, 256)), dim3(256), 0, (cudaStream_t)stream, (const float*)x, out, a, r, c, accumulate);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_bcast_mid(const float* x, void* out, int32_t a, int32_t r, int32_t c, void* stream) {
  TOIST_REQUIRE(x && out, "toist_bcast_mid: null pointer");
  const long long n = (long long)a * r * c;
  if (n == 0) return TOIST_OK;
  launch_pdl(bcast_mid_kernel, dim3(nblk(n, 256)), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)out, a, r, c);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_nchw_to_nhwc(const float* x, void* y, int32_t n, int32_t c, int32_t hw, void* stream) {
  TOIST_REQUIRE(x && y, "toist_nchw_to_nhwc: null pointer");
  const long long t = (long long)n * c * hw;
  if (t == 0) return TOIST_OK;
  launch_pdl(nchw_to_nhwc_kernel, dim3(nblk(t, 256)), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, n, c, hw);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_nhwc_to_nchw(const void* x, float* y, int32_t n, int32_t c, int32_t hw, void* stream) {
  TOIST_REQUIRE(x && y, "toist_nhwc_to_nchw: null pointer");
  const long long t = (long long)n * c * hw;
  if (t == 0) return TOIST_OK;
  launch_pdl(nhwc_to_nchw_kernel, dim3(nblk(t, 256)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, y, n, c, hw);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_cast_pad_f32_bf16(const float* src, void* dst, int64_t rows, int32_t n, int32_t ld, void* stream) {
  TOIST_REQUIRE(src && dst && ld >= n, "toist_cast_pad_f32_bf16: bad arguments");
  if (rows * ld == 0) return TOIST_OK;
  launch_pdl(cast_pad_kernel, dim3(nblk(rows * ld, 256)), dim3(256), 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst, rows, n, ld);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_dropout(const void* x, const void* res, void* out, int64_t n, int32_t dtype, float p, const uint64_t* seed,
                  uint32_t site, void* stream) {
  TOIST_REQUIRE(x && out && seed, "toist_dropout: null pointer");
  TOIST_REQUIRE(p >= 0.f && p < 1.f, "toist_dropout: p = %f out of [0, 1)", p);
  if (n == 0) return TOIST_OK;
  const uint32_t thr = (uint32_t)((double)p * 4294967296.0);
  const float scale = 1.f / (1.f - p);
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(out);
  if (dtype == TOIST_BF16 && n % 8 == 0 && (al & 15) == 0)
    launch_pdl(dropout_bf16_vec_kernel, dim3(nblk(n / 8, 256)), dim3(256), 0, (cudaStream_t)stream,
               (const __nv_bfloat16*)x, (const __nv_bfloat16*)res, (__nv_bfloat16*)out, n / 8,
               (const unsigned long long*)seed, site, thr, scale);
  else if (dtype == TOIST_BF16)
    launch_pdl((dropout_kernel<__nv_bfloat16>), dim3(nblk(n, 256)), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)res, (__nv_bfloat16*)out, n, (const unsigned long long*)seed, site, thr, scale);
  else
    launch_pdl((dropout_kernel<float>), dim3(nblk(n, 256)), dim3(256), 0, (cudaStream_t)stream, 
        (const float*)x, (const float*)res, (float*)out, n, (const unsigned long long*)seed, site, thr, scale);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_permute_021(const float* src, float* dst, int32_t a, int32_t b, int32_t c, void* stream) {
  TOIST_REQUIRE(src && dst, "toist_permute_021: null pointer");
  const long long t = (long long)a * b * c;
  if (t == 0) return TOIST_OK;
  launch_pdl(permute_021_kernel, dim3(nblk(t, 256)), dim3(256), 0, (cudaStream_t)stream, src, dst, a, b, c);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_key_mask(const uint8_t* pad_mask, const int64_t* text_attention, uint8_t* small_mask, uint8_t* key_mask,
                   int32_t batch, int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w, int32_t n_text,
                   void* stream) {
  TOIST_REQUIRE(pad_mask && (small_mask || key_mask), "toist_key_mask: null pointer");
  TOIST_REQUIRE(n_text == 0 || text_attention != nullptr || key_mask == nullptr, "toist_key_mask: text mask missing");
  const long long t = (long long)batch * (out_h * out_w + n_text);
  if (t == 0) return TOIST_OK;
  launch_pdl(key_mask_kernel, dim3(nblk(t, 256)), dim3(256), 0, (cudaStream_t)stream, pad_mask, (const long long*)text_attention,
                                                                  small_mask, key_mask, batch, in_h, in_w, out_h,
                                                                  out_w, n_text);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

}  // extern "C"
