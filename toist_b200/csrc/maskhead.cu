// Mask branch of DETRsegm (reference models/segmentation.py:157-241,244-273; models/mdetr.py:827-853):
// input assembly for MaskHeadSmallConv, GroupNorm(+ReLU) forward / backward on NHWC bf16 maps, nearest upsample +
// FPN add and its backward, and the fused bilinear-upsample + focal + dice mask losses.  The 3x3 / 1x1 convolutions
// themselves run on the implicit-GEMM engine (gemm.cu).
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace toist {

// ------------------------------------------------------------------------------------------------ head input
// x0[(b*Q + q), pix, c] = c < E ? src_proj[pix, b, c] : attn[b, c - E, q, pix]      (segmentation.py:205-207)
// src_proj: bf16 sequence layout [HW, B, E]; attn: bf16 [B, NH, Q, ld] (softmax over pix per head); x0 NHWC bf16.
__global__ void mask_input_kernel(const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ attn,
                                  __nv_bfloat16* __restrict__ x0, int B, int Q, int HW, int E, int NH, int ld) {
  pdl_prologue();
  const int C = E + NH;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * Q * HW * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  long long t = i / C;
  const int pix = (int)(t % HW);
  t /= HW;
  const int q = (int)(t % Q);
  const int b = (int)(t / Q);
  x0[i] = c < E ? src[((long long)pix * B + b) * E + c] : attn[(((long long)b * NH + (c - E)) * Q + q) * ld + pix];
}

// backward: dsrc[pix, b, c] = sum_q dx0[(b*Q+q), pix, c];  dattn[b, h, q, pix] (fp32, ld_s) = dx0[(b*Q+q), pix, E + h]
__global__ void mask_input_bwd_kernel(const __nv_bfloat16* __restrict__ dx0, __nv_bfloat16* __restrict__ dsrc,
                                      float* __restrict__ dattn, int B, int Q, int HW, int E, int NH, int ld) {
  pdl_prologue();
  const int C = E + NH;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * HW * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  long long t = i / C;
  const int pix = (int)(t % HW);
  const int b = (int)(t / HW);
  if (c < E) {
    if (dsrc == nullptr) return;
    float acc = 0.f;
    for (int q = 0; q < Q; ++q) acc += __bfloat162float(dx0[(((long long)b * Q + q) * HW + pix) * C + c]);
    dsrc[((long long)pix * B + b) * E + c] = __float2bfloat16_rn(acc);
  } else {
    const int h = c - E;
    for (int q = 0; q < Q; ++q)
      dattn[(((long long)b * NH + h) * Q + q) * ld + pix] =
          __bfloat162float(dx0[(((long long)b * Q + q) * HW + pix) * C + c]);
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm + ReLU
// z NHWC bf16 [N, HW, C]; one CTA per map.  stats[n, g] = (mean, rstd) over the C/G channels x HW pixels of group g.
constexpr int kGNMaxC = 512;
constexpr int kGNMaxThreads = 256;

__global__ void groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ z, float* __restrict__ mean,
                                       float* __restrict__ rstd, int HW, int C, int G, float eps) {
  pdl_prologue();
  // deterministic: per-thread partials -> fixed-order sum per channel -> fixed-order sum per group (no atomics)
  __shared__ float s_sum[kGNMaxC], s_sq[kGNMaxC];
  __shared__ float p_sum[kGNMaxThreads * 8], p_sq[kGNMaxThreads * 8];
  const int n = blockIdx.x;
  const int chunks = C / 8;                 // 16-byte chunks per pixel
  const int chunk = threadIdx.x % chunks;   // fixed channel chunk per thread (blockDim is a multiple of `chunks`)
  const int prow = threadIdx.x / chunks, pstride = blockDim.x / chunks;
  const __nv_bfloat16* zn = z + (long long)n * HW * C;
  float a[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = q[k] = 0.f;
  for (int p = prow; p < HW; p += pstride) {
    const uint4 u = *reinterpret_cast<const uint4*>(zn + (long long)p * C + chunk * 8);
    const uint32_t* pu = &u.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(pu[k]);
      a[2 * k] += f.x; q[2 * k] += f.x * f.x;
      a[2 * k + 1] += f.y; q[2 * k + 1] += f.y * f.y;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    p_sum[(prow * chunks + chunk) * 8 + k] = a[k];   // == [prow][channel]
    p_sq[(prow * chunks + chunk) * 8 + k] = q[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, ss = 0.f;
    for (int r = 0; r < pstride; ++r) {
      s += p_sum[r * C + c];
      ss += p_sq[r * C + c];
    }
    s_sum[c] = s;
    s_sq[c] = ss;
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f, ss = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      s += s_sum[c];
      ss += s_sq[c];
    }
    const float m = s / ((float)cpg * HW);
    const float var = fmaxf(ss / ((float)cpg * HW) - m * m, 0.f);
    mean[n * G + g] = m;
    rstd[n * G + g] = rsqrtf(var + eps);
  }
}

// a = relu((z - mean) * rstd * gamma + beta), 8 channels per thread
__global__ void groupnorm_relu_fwd_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                                          const float* __restrict__ rstd, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, __nv_bfloat16* __restrict__ a, long long total8,
                                          int HW, int C, int G) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int chunks = C / 8;
  const int c0 = (int)(i % chunks) * 8;
  const long long n = i / ((long long)chunks * HW);
  const int cpg = C / G;
  const uint4 u = *reinterpret_cast<const uint4*>(z + i * 8);
  const uint32_t* pu = &u.x;
  float v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = unpack_bf16(pu[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k, g = c / cpg;
    const float y = (v[k] - mean[n * G + g]) * rstd[n * G + g] * gamma[c] + beta[c];
    v[k] = fmaxf(y, 0.f);
  }
  uint4 o;
  o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(a + i * 8) = o;
}

// backward pass 1 (one CTA per map): dy = da * (y > 0); per group s1 = sum dy*gamma, s2 = sum dy*gamma*xhat;
// per channel dgamma += sum dy*xhat, dbeta += sum dy (global atomics, once per CTA and channel)
__global__ void groupnorm_relu_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ z,
                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 float* __restrict__ s1, float* __restrict__ s2,
                                                 float* __restrict__ dgamma, float* __restrict__ dbeta, int HW, int C,
                                                 int G) {
  pdl_prologue();
  __shared__ float c_dyx[kGNMaxC], c_dy[kGNMaxC];
  __shared__ float p_dyx[kGNMaxThreads * 8], p_dy[kGNMaxThreads * 8];
  const int n = blockIdx.x;
  const int chunks = C / 8, cpg = C / G;
  const int chunk = threadIdx.x % chunks;
  const int prow = threadIdx.x / chunks, pstride = blockDim.x / chunks;
  const long long base = (long long)n * HW * C;
  float adyx[8], ady[8], mu[8], rs[8], ga[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = chunk * 8 + k, g = c / cpg;
    adyx[k] = ady[k] = 0.f;
    mu[k] = mean[n * G + g];
    rs[k] = rstd[n * G + g];
    ga[k] = gamma[c];
    be[k] = beta[c];
  }
  {
    for (int p = prow; p < HW; p += pstride) {
      const long long off = base + (long long)p * C + chunk * 8;
      const uint4 uz = *reinterpret_cast<const uint4*>(z + off), ud = *reinterpret_cast<const uint4*>(da + off);
      const uint32_t* pz = &uz.x;
      const uint32_t* pd = &ud.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fz = unpack_bf16(pz[k]), fd = unpack_bf16(pd[k]);
        const float zz[2] = {fz.x, fz.y}, dd[2] = {fd.x, fd.y};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int kk = 2 * k + j;
          const float xh = (zz[j] - mu[kk]) * rs[kk];
          const float dy = (xh * ga[kk] + be[kk] > 0.f) ? dd[j] : 0.f;
          adyx[kk] += dy * xh;
          ady[kk] += dy;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      p_dyx[(prow * chunks + chunk) * 8 + k] = adyx[k];
      p_dy[(prow * chunks + chunk) * 8 + k] = ady[k];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int r = 0; r < pstride; ++r) {
      a += p_dyx[r * C + c];
      b += p_dy[r * C + c];
    }
    c_dyx[c] = a;
    c_dy[c] = b;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += c_dy[c] * gamma[c];
      b += c_dyx[c] * gamma[c];
    }
    s1[n * G + g] = a;
    s2[n * G + g] = b;
  }
  if (dgamma != nullptr)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      atomicAdd(dgamma + c, c_dyx[c]);
      atomicAdd(dbeta + c, c_dy[c]);
    }
}

// backward pass 2: dz = rstd * (dy*gamma - (s1 + xhat*s2) / M),  M = (C/G) * HW
__global__ void groupnorm_relu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ z,
                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* __restrict__ s1, const float* __restrict__ s2,
                                                __nv_bfloat16* __restrict__ dz, long long total8, int HW, int C, int G) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int chunks = C / 8, cpg = C / G;
  const int c0 = (int)(i % chunks) * 8;
  const long long n = i / ((long long)chunks * HW);
  const float invM = 1.f / ((float)cpg * HW);
  const uint4 uz = *reinterpret_cast<const uint4*>(z + i * 8), ud = *reinterpret_cast<const uint4*>(da + i * 8);
  const uint32_t* pz = &uz.x;
  const uint32_t* pd = &ud.x;
  float o[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 fz = unpack_bf16(pz[k]), fd = unpack_bf16(pd[k]);
    const float zz[2] = {fz.x, fz.y}, dd[2] = {fd.x, fd.y};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = c0 + 2 * k + j, g = c / cpg;
      const float rs = rstd[n * G + g];
      const float xh = (zz[j] - mean[n * G + g]) * rs;
      const float dy = (xh * gamma[c] + beta[c] > 0.f) ? dd[j] : 0.f;
      o[2 * k + j] = rs * (dy * gamma[c] - (s1[n * G + g] + xh * s2[n * G + g]) * invM);
    }
  }
  uint4 u;
  u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
  *reinterpret_cast<uint4*>(dz + i * 8) = u;
}

// ------------------------------------------------------------------------------------------------ upsample + FPN add
// out[n, y, x, c] = fpn[n / Q, y, x, c] + xs[n, floor(y*h/H), floor(x*w/W), c]     (segmentation.py:217-220)
__global__ void upsample_add_kernel(const __nv_bfloat16* __restrict__ xs, const __nv_bfloat16* __restrict__ fpn,
                                    __nv_bfloat16* __restrict__ out, long long total8, int Q, int H, int W, int h, int w,
                                    int C) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int chunks = C / 8;
  const int ch = (int)(i % chunks);
  long long t = i / chunks;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H);
  const long long n = t / H;
  int sy = (int)floorf((float)y * ((float)h / (float)H)), sx = (int)floorf((float)x * ((float)w / (float)W));
  sy = sy < h - 1 ? sy : h - 1;
  sx = sx < w - 1 ? sx : w - 1;
  const uint4 a = *reinterpret_cast<const uint4*>(xs + (((n * h + sy) * w + sx) * C) + ch * 8);
  const uint4 f = *reinterpret_cast<const uint4*>(fpn + ((((n / Q) * H + y) * W + x) * C) + ch * 8);
  const uint32_t* pa = &a.x;
  const uint32_t* pf = &f.x;
  uint4 o;
  uint32_t* po = &o.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 fa = unpack_bf16(pa[k]), ff = unpack_bf16(pf[k]);
    po[k] = pack_bf16(fa.x + ff.x, fa.y + ff.y);
  }
  *reinterpret_cast<uint4*>(out + i * 8) = o;
}

// dxs[n, sy, sx, c] = sum of dout over the destination pixels that read (sy, sx)
__global__ void upsample_bwd_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dxs,
                                    long long total8, int H, int W, int h, int w, int C) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int chunks = C / 8;
  const int ch = (int)(i % chunks);
  long long t = i / chunks;
  const int sx = (int)(t % w); t /= w;
  const int sy = (int)(t % h);
  const long long n = t / h;
  // destination rows / columns mapping to this source pixel: the smallest y with floor(y*h/H) >= sy, etc.
  int y0 = (int)ceilf((float)sy * (float)H / (float)h), y1 = (int)ceilf((float)(sy + 1) * (float)H / (float)h);
  int x0 = (int)ceilf((float)sx * (float)W / (float)w), x1 = (int)ceilf((float)(sx + 1) * (float)W / (float)w);
  y0 = max(y0 - 1, 0); y1 = min(y1 + 1, H); x0 = max(x0 - 1, 0); x1 = min(x1 + 1, W);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int y = y0; y < y1; ++y) {
    int ty = (int)floorf((float)y * ((float)h / (float)H));
    ty = ty < h - 1 ? ty : h - 1;
    if (ty != sy) continue;
    for (int x = x0; x < x1; ++x) {
      int tx = (int)floorf((float)x * ((float)w / (float)W));
      tx = tx < w - 1 ? tx : w - 1;
      if (tx != sx) continue;
      const uint4 u = *reinterpret_cast<const uint4*>(dout + (((n * H + y) * W + x) * C) + ch * 8);
      const uint32_t* pu = &u.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16(pu[k]);
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
  }
  uint4 o;
  o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]); o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(dxs + i * 8) = o;
}

// dfpn[b, pix, c] = sum_q dout[b*Q + q, pix, c]   (the adapter output is shared by the Q queries of an image)
__global__ void sum_queries_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dfpn,
                                   long long per_image8, int Q) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = blockIdx.y;
  if (i >= per_image8) return;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int q = 0; q < Q; ++q) {
    const uint4 u = *reinterpret_cast<const uint4*>(dout + ((b * Q + q) * per_image8 + i) * 8);
    const uint32_t* pu = &u.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(pu[k]);
      acc[2 * k] += f.x;
      acc[2 * k + 1] += f.y;
    }
  }
  uint4 o;
  o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]); o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(dfpn + (b * per_image8 + i) * 8) = o;
}

// ------------------------------------------------------------------------------------------------ mask losses
// For every matched (image b, target t) pair: bilinear upsample (align_corners = false) of pred_masks[b, q] from
// (hm, wm) to (HT, WT), sigmoid, focal (alpha .25, gamma 2) and dice partial sums against the padded target mask.
// pred fp32 [B, Q, hm, wm]; tgt uint8 [B, Tmax, HT, WT]; match_q int32 [B, Tmax] (query of target t, -1 = none).
// sums[b, t, 4] = { sum focal, sum p*t, sum p, sum t }.   grid = (chunks, B*Tmax)
__device__ __forceinline__ void bilinear_src(int d, int in, int out, int& i0, int& i1, float& l1) {
  float s = ((float)d + 0.5f) * ((float)in / (float)out) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i0 = i0 < in - 1 ? i0 : in - 1;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  l1 = s - (float)i0;
}

__global__ void mask_loss_fwd_kernel(const float* __restrict__ pred, const uint8_t* __restrict__ tgt,
                                     const int* __restrict__ match_q, const int* __restrict__ tgt_count,
                                     float* __restrict__ sums, int Q, int Tmax, int hm, int wm, int HT, int WT) {
  pdl_prologue();
  const int pair = blockIdx.y;
  const int b = pair / Tmax, t = pair % Tmax;
  if (t >= min(tgt_count[b], Tmax)) return;
  const int q = match_q[pair];
  if (q < 0) return;
  const float* pm = pred + ((long long)b * Q + q) * hm * wm;
  const uint8_t* tm = tgt + (long long)pair * HT * WT;
  float f = 0.f, pt = 0.f, ps = 0.f, ts = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)HT * WT;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / WT), x = (int)(i % WT);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(y, hm, HT, y0, y1, ly);
    bilinear_src(x, wm, WT, x0, x1, lx);
    const float v = (1.f - ly) * ((1.f - lx) * pm[y0 * wm + x0] + lx * pm[y0 * wm + x1]) +
                    ly * ((1.f - lx) * pm[y1 * wm + x0] + lx * pm[y1 * wm + x1]);
    const float tt = tm[i] ? 1.f : 0.f;
    const float p = 1.f / (1.f + expf(-v));
    // binary_cross_entropy_with_logits: max(v,0) - v*t + log1p(exp(-|v|))
    const float ce = fmaxf(v, 0.f) - v * tt + log1pf(expf(-fabsf(v)));
    const float p_t = p * tt + (1.f - p) * (1.f - tt);
    const float a_t = 0.25f * tt + 0.75f * (1.f - tt);
    f += a_t * ce * (1.f - p_t) * (1.f - p_t);
    pt += p * tt;
    ps += p;
    ts += tt;
  }
  f = warp_sum(f); pt = warp_sum(pt); ps = warp_sum(ps); ts = warp_sum(ts);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums + pair * 4 + 0, f);
    atomicAdd(sums + pair * 4 + 1, pt);
    atomicAdd(sums + pair * 4 + 2, ps);
    atomicAdd(sums + pair * 4 + 3, ts);
  }
}

// dpred[b, q, :, :] += g_focal * dfocal/dv / (HT*WT*num_boxes) + g_dice * ddice/dv / num_boxes, scattered through the
// bilinear weights (atomics; dpred pre-zeroed).  dice = 1 - (2*PT + 1) / (PS + TS + 1).
__global__ void mask_loss_bwd_kernel(const float* __restrict__ pred, const uint8_t* __restrict__ tgt,
                                     const int* __restrict__ match_q, const int* __restrict__ tgt_count,
                                     const float* __restrict__ sums, const float* __restrict__ num_boxes,
                                     const float* __restrict__ gout /* [2]: d/d loss_mask, d/d loss_dice */,
                                     float* __restrict__ dpred, int Q, int Tmax, int hm, int wm, int HT, int WT) {
  pdl_prologue();
  const int pair = blockIdx.y;
  const int b = pair / Tmax, t = pair % Tmax;
  if (t >= min(tgt_count[b], Tmax)) return;
  const int q = match_q[pair];
  if (q < 0) return;
  const float* pm = pred + ((long long)b * Q + q) * hm * wm;
  float* dm = dpred + ((long long)b * Q + q) * hm * wm;
  const uint8_t* tm = tgt + (long long)pair * HT * WT;
  const float inv_nb = 1.f / num_boxes[0];
  const float gf = gout[0] * inv_nb / ((float)HT * (float)WT), gd = gout[1] * inv_nb;
  const float PT = sums[pair * 4 + 1], den = sums[pair * 4 + 2] + sums[pair * 4 + 3] + 1.f;
  const float num = 2.f * PT + 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)HT * WT;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / WT), x = (int)(i % WT);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(y, hm, HT, y0, y1, ly);
    bilinear_src(x, wm, WT, x0, x1, lx);
    const float v = (1.f - ly) * ((1.f - lx) * pm[y0 * wm + x0] + lx * pm[y0 * wm + x1]) +
                    ly * ((1.f - lx) * pm[y1 * wm + x0] + lx * pm[y1 * wm + x1]);
    const float tt = tm[i] ? 1.f : 0.f;
    const float p = 1.f / (1.f + expf(-v));
    const float ce = fmaxf(v, 0.f) - v * tt + log1pf(expf(-fabsf(v)));
    const float p_t = p * tt + (1.f - p) * (1.f - tt);
    const float a_t = 0.25f * tt + 0.75f * (1.f - tt);
    const float om = 1.f - p_t;
    // d ce / dv = p - t ; d p_t / dv = (2t - 1) * p * (1 - p)
    const float dfocal = a_t * ((p - tt) * om * om - ce * 2.f * om * (2.f * tt - 1.f) * p * (1.f - p));
    // d dice / dp = -(2 t * den - num) / den^2
    const float ddice = -(2.f * tt * den - num) / (den * den) * p * (1.f - p);
    const float g = gf * dfocal + gd * ddice;
    atomicAdd(dm + y0 * wm + x0, g * (1.f - ly) * (1.f - lx));
    atomicAdd(dm + y0 * wm + x1, g * (1.f - ly) * lx);
    atomicAdd(dm + y1 * wm + x0, g * ly * (1.f - lx));
    atomicAdd(dm + y1 * wm + x1, g * ly * lx);
  }
}

// out[0] = loss_mask = sum_pairs (focal_sum / (HT*WT)) / num_boxes ; out[1] = loss_dice
__global__ void mask_loss_reduce_kernel(const float* __restrict__ sums, const int* __restrict__ match_q,
                                        const int* __restrict__ tgt_count, const float* __restrict__ num_boxes,
                                        float* __restrict__ out, int B, int Tmax, float inv_pix) {
  pdl_prologue();
  float lm = 0.f, ld = 0.f;
  for (int pair = threadIdx.x; pair < B * Tmax; pair += blockDim.x) {
    const int b = pair / Tmax, t = pair % Tmax;
    if (t >= min(tgt_count[b], Tmax) || match_q[pair] < 0) continue;
    lm += sums[pair * 4 + 0] * inv_pix;
    ld += 1.f - (2.f * sums[pair * 4 + 1] + 1.f) / (sums[pair * 4 + 2] + sums[pair * 4 + 3] + 1.f);
  }
  __shared__ float ra[32], rb[32];
  lm = warp_sum(lm);
  ld = warp_sum(ld);
  if ((threadIdx.x & 31) == 0) {
    ra[threadIdx.x >> 5] = lm;
    rb[threadIdx.x >> 5] = ld;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b2 = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      a += ra[i];
      b2 += rb[i];
    }
    out[0] = a / num_boxes[0];
    out[1] = b2 / num_boxes[0];
  }
}

}  // namespace toist

using namespace toist;

static inline unsigned nblk(long long n, int per) { return (unsigned)((n + per - 1) / per); }

// threads per CTA for the per-map GroupNorm reductions: a multiple of C/8 close to 256 (each thread owns one chunk)
static int gn_threads(int C) {
  const int chunks = C / 8;
  int t = (kGNMaxThreads / chunks) * chunks;  // C <= 512 -> chunks <= 64 <= kGNMaxThreads
  return t;
}

extern "C" {

int toist_mask_input(const void* src_proj, const void* attn, void* x0, int32_t batch, int32_t n_queries, int32_t hw,
                     int32_t dim, int32_t n_heads, int32_t ld_attn, void* stream) {
  TOIST_REQUIRE(src_proj && attn && x0, "toist_mask_input: null pointer");
  const long long total = (long long)batch * n_queries * hw * (dim + n_heads);
  if (total == 0) return TOIST_OK;
  launch_pdl(mask_input_kernel, dim3(nblk(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)src_proj, (const __nv_bfloat16*)attn, (__nv_bfloat16*)x0, batch, n_queries, hw, dim, n_heads,
      ld_attn);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_mask_input_bwd(const void* dx0, void* dsrc_proj, float* dattn, int32_t batch, int32_t n_queries, int32_t hw,
                         int32_t dim, int32_t n_heads, int32_t ld_attn, void* stream) {
  TOIST_REQUIRE(dx0 && dattn, "toist_mask_input_bwd: null pointer");
  const long long total = (long long)batch * hw * (dim + n_heads);
  if (total == 0) return TOIST_OK;
  launch_pdl(mask_input_bwd_kernel, dim3(nblk(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)dx0, (__nv_bfloat16*)dsrc_proj, dattn, batch, n_queries, hw, dim, n_heads, ld_attn);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_groupnorm_relu_fwd(const void* z, const float* gamma, const float* beta, void* a, float* mean, float* rstd,
                             int32_t n_maps, int32_t hw, int32_t channels, int32_t groups, float eps, void* stream) {
  TOIST_REQUIRE(z && gamma && beta && a && mean && rstd, "toist_groupnorm_relu_fwd: null pointer");
  TOIST_REQUIRE(channels % 8 == 0 && channels <= kGNMaxC && groups >= 1 && channels % groups == 0,
                "toist_groupnorm_relu_fwd: %d channels / %d groups unsupported", channels, groups);
  if (n_maps == 0) return TOIST_OK;
  cudaStream_t st = (cudaStream_t)stream;
  launch_pdl(groupnorm_stats_kernel, dim3(n_maps), dim3(gn_threads(channels)), 0, st, (const __nv_bfloat16*)z, mean, rstd, hw, channels,
                                                                  groups, eps);
  const long long total8 = (long long)n_maps * hw * (channels / 8);
  launch_pdl(groupnorm_relu_fwd_kernel, dim3(nblk(total8, 256)), dim3(256), 0, st, (const __nv_bfloat16*)z, mean, rstd, gamma, beta,
                                                               (__nv_bfloat16*)a, total8, hw, channels, groups);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_groupnorm_relu_bwd(const void* da, const void* z, const float* mean, const float* rstd, const float* gamma,
                             const float* beta, void* dz, float* dgamma, float* dbeta, float* scratch, int32_t n_maps,
                             int32_t hw, int32_t channels, int32_t groups, void* stream) {
  TOIST_REQUIRE(da && z && mean && rstd && gamma && beta && dz && scratch, "toist_groupnorm_relu_bwd: null pointer");
  TOIST_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "toist_groupnorm_relu_bwd: pass both dgamma and dbeta");
  TOIST_REQUIRE(channels % 8 == 0 && channels <= kGNMaxC && channels % groups == 0, "toist_groupnorm_relu_bwd: bad shape");
  if (n_maps == 0) return TOIST_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* s1 = scratch;
  float* s2 = scratch + (size_t)n_maps * groups;
  launch_pdl(groupnorm_relu_bwd_reduce_kernel, dim3(n_maps), dim3(gn_threads(channels)), 0, st, 
      (const __nv_bfloat16*)da, (const __nv_bfloat16*)z, mean, rstd, gamma, beta, s1, s2, dgamma, dbeta, hw, channels,
      groups);
  const long long total8 = (long long)n_maps * hw * (channels / 8);
  launch_pdl(groupnorm_relu_bwd_apply_kernel, dim3(nblk(total8, 256)), dim3(256), 0, st, 
      (const __nv_bfloat16*)da, (const __nv_bfloat16*)z, mean, rstd, gamma, beta, s1, s2, (__nv_bfloat16*)dz, total8, hw,
      channels, groups);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_upsample_add(const void* xs, const void* fpn, void* out, int32_t n_maps, int32_t n_queries, int32_t out_h,
                       int32_t out_w, int32_t in_h, int32_t in_w, int32_t channels, void* stream) {
  TOIST_REQUIRE(xs && fpn && out && channels % 8 == 0 && n_queries >= 1, "toist_upsample_add: bad arguments");
  const long long total8 = (long long)n_maps * out_h * out_w * (channels / 8);
  if (total8 == 0) return TOIST_OK;
  launch_pdl(upsample_add_kernel, dim3(nblk(total8, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)xs, (const __nv_bfloat16*)fpn, (__nv_bfloat16*)out, total8, n_queries, out_h, out_w, in_h,
      in_w, channels);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_upsample_add_bwd(const void* dout, void* dxs, void* dfpn, int32_t n_maps, int32_t n_queries, int32_t out_h,
                           int32_t out_w, int32_t in_h, int32_t in_w, int32_t channels, void* stream) {
  TOIST_REQUIRE(dout && dxs && channels % 8 == 0 && n_queries >= 1, "toist_upsample_add_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long t8 = (long long)n_maps * in_h * in_w * (channels / 8);
  if (t8 == 0) return TOIST_OK;
  launch_pdl(upsample_bwd_kernel, dim3(nblk(t8, 256)), dim3(256), 0, st, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dxs, t8, out_h, out_w,
                                                     in_h, in_w, channels);
  if (dfpn != nullptr) {
    const long long per8 = (long long)out_h * out_w * (channels / 8);
    dim3 grid(nblk(per8, 256), n_maps / n_queries);
    launch_pdl(sum_queries_kernel, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dfpn, per8, n_queries);
  }
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_mask_loss_fwd(const float* pred_masks, const uint8_t* tgt_masks, const int32_t* match_q,
                        const int32_t* tgt_count, const float* num_boxes, float* sums, float* out, int32_t batch,
                        int32_t n_queries, int32_t t_max, int32_t mask_h, int32_t mask_w, int32_t tgt_h, int32_t tgt_w,
                        void* stream) {
  TOIST_REQUIRE(pred_masks && tgt_masks && match_q && tgt_count && num_boxes && sums && out,
                "toist_mask_loss_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  TOIST_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 4 * batch * t_max, st));
  const long long pix = (long long)tgt_h * tgt_w;
  int chunks = (int)((pix + 256 * 8 - 1) / (256 * 8));
  if (chunks > 64) chunks = 64;
  dim3 grid(chunks, batch * t_max);
  launch_pdl(mask_loss_fwd_kernel, dim3(grid), dim3(256), 0, st, pred_masks, tgt_masks, match_q, tgt_count, sums, n_queries, t_max, mask_h,
                                             mask_w, tgt_h, tgt_w);
  launch_pdl(mask_loss_reduce_kernel, dim3(1), dim3(256), 0, st, sums, match_q, tgt_count, num_boxes, out, batch, t_max,
                                             1.f / ((float)tgt_h * (float)tgt_w));
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_mask_loss_bwd(const float* pred_masks, const uint8_t* tgt_masks, const int32_t* match_q,
                        const int32_t* tgt_count, const float* sums, const float* num_boxes, const float* gout,
                        float* dpred, int32_t batch, int32_t n_queries, int32_t t_max, int32_t mask_h, int32_t mask_w,
                        int32_t tgt_h, int32_t tgt_w, void* stream) {
  TOIST_REQUIRE(pred_masks && tgt_masks && match_q && tgt_count && sums && num_boxes && gout && dpred,
                "toist_mask_loss_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  TOIST_CHECK_CUDA(cudaMemsetAsync(dpred, 0, sizeof(float) * (size_t)batch * n_queries * mask_h * mask_w, st));
  const long long pix = (long long)tgt_h * tgt_w;
  int chunks = (int)((pix + 256 * 8 - 1) / (256 * 8));
  if (chunks > 64) chunks = 64;
  dim3 grid(chunks, batch * t_max);
  launch_pdl(mask_loss_bwd_kernel, dim3(grid), dim3(256), 0, st, pred_masks, tgt_masks, match_q, tgt_count, sums, num_boxes, gout, dpred,
                                             n_queries, t_max, mask_h, mask_w, tgt_h, tgt_w);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

}  // extern "C"
