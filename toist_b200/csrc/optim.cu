// Optimizer side of the training step, multi-tensor and HBM-bound: every kernel walks a device table
// of (pointer, length) records, one CTA per 2048-element chunk, 16-byte accesses, so the ~950 parameter tensors of the
// model cost one launch instead of one (or several) launches each.
//   toist_grad_sqnorm + toist_grad_clip_scale : torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)  (engine.py:89-90)
//   toist_adamw_step                          : torch.optim.AdamW(param_dicts).step()                      (main.py:351-392, engine.py:91)
//   toist_ema_update                          : util/optim.py:9-26 update_ema (w_ema = w_ema * decay + (1 - decay) * w)
// Algorithmic bytes per parameter: clip 12 B (read g twice, write once), AdamW 28 B (read p, g, m, v; write p, m, v),
// EMA 12 B; 185 M parameters -> 9.6 GB -> 1.5 ms at the measured 6.5 TB/s.
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace toist {

constexpr int kOptChunk = 2048;  // elements per CTA (256 threads x 8)

struct OptItem {  // 48 bytes, mirrored by toist_b200/util/optim.py
  float* a;       // clip: grad | adamw: param | ema: ema tensor
  float* b;       // adamw: grad | ema: model tensor
  float* c;       // adamw: exp_avg
  float* d;       // adamw: exp_avg_sq
  long long n;
  int first_block;
  int group;
};

__device__ __forceinline__ const OptItem& find_item(const OptItem* items, int n_items, int block) {
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= block) lo = mid; else hi = mid - 1;
  }
  return items[lo];
}

// partial[block] = sum of squares of this block's chunk (fixed summation order: deterministic)
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const OptItem* __restrict__ items, int n_items,
                                                          float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float red[8];
  const OptItem it = find_item(items, n_items, blockIdx.x);
  const long long i = (long long)(blockIdx.x - it.first_block) * kOptChunk + (long long)threadIdx.x * 8;
  float s = 0.f;
  if (i + 8 <= it.n && (reinterpret_cast<uintptr_t>(it.a) & 15) == 0) {
    const float4 x = *reinterpret_cast<const float4*>(it.a + i), y = *reinterpret_cast<const float4*>(it.a + i + 4);
    s = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w + y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
  } else {
    for (long long k = i; k < it.n && k < i + 8; ++k) s += it.a[k] * it.a[k];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

// out[0] = total L2 norm, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6))   (one CTA, fixed order)
__global__ void __launch_bounds__(1024) grad_norm_finish_kernel(const float* __restrict__ partial, int n, float max_norm,
                                                                float* __restrict__ out) {
  pdl_prologue();
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    const float norm = (float)sqrt(t);
    out[0] = norm;
    out[1] = fminf(1.f, max_norm / (norm + 1e-6f));
  }
}

// g *= coef[1]  (in place, like clip_grad_norm_)
__global__ void __launch_bounds__(256) grad_scale_kernel(const OptItem* __restrict__ items, int n_items,
                                                         const float* __restrict__ coef) {
  pdl_prologue();
  const float c = coef[1];
  if (c >= 1.f) return;  // torch multiplies by the clamped coefficient 1.0: a no-op
  const OptItem it = find_item(items, n_items, blockIdx.x);
  const long long i = (long long)(blockIdx.x - it.first_block) * kOptChunk + (long long)threadIdx.x * 8;
  if (i + 8 <= it.n && (reinterpret_cast<uintptr_t>(it.a) & 15) == 0) {
    float4* p = reinterpret_cast<float4*>(it.a + i);
    float4 x = p[0], y = p[1];
    x.x *= c; x.y *= c; x.z *= c; x.w *= c; y.x *= c; y.y *= c; y.z *= c; y.w *= c;
    p[0] = x; p[1] = y;
  } else {
    for (long long k = i; k < it.n && k < i + 8; ++k) it.a[k] *= c;
  }
}

// Scalars are prepared on the host in double precision exactly as torch's Python code forms them (1 - beta2 evaluated
// in fp32 is off by 5e-5 relative, which shows in exp_avg_sq), then rounded to fp32 once.
struct AdamHyper {
  float decay;        // 1 - lr * weight_decay
  float omb1;         // 1 - beta1
  float beta2;
  float omb2;         // 1 - beta2
  float eps;
  float step_size;    // lr / (1 - beta1^t)
  float bc2_sqrt;     // sqrt(1 - beta2^t)
  float pad;
};
constexpr int kMaxAdamRows = 32;  // (parameter group, step count) rows per launch: 6 groups in the distillation recipe
struct AdamGroups {
  AdamHyper g[kMaxAdamRows];
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamHyper& h) {
  // torch.optim.AdamW (_single_tensor_adamw): decoupled decay, lerp, mul + addcmul, sqrt / bias correction + eps, addcdiv
  p *= h.decay;
  m += (g - m) * h.omb1;
  v = v * h.beta2 + h.omb2 * g * g;
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p -= h.step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_kernel(const OptItem* __restrict__ items, int n_items,
                                                    const __grid_constant__ AdamGroups groups) {
  pdl_prologue();
  const OptItem it = find_item(items, n_items, blockIdx.x);
  const AdamHyper h = groups.g[it.group];
  const long long i = (long long)(blockIdx.x - it.first_block) * kOptChunk + (long long)threadIdx.x * 8;
  const uintptr_t al = reinterpret_cast<uintptr_t>(it.a) | reinterpret_cast<uintptr_t>(it.b) |
                       reinterpret_cast<uintptr_t>(it.c) | reinterpret_cast<uintptr_t>(it.d);
  if (i + 8 <= it.n && (al & 15) == 0) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const long long j = i + half * 4;
      float4 p = *reinterpret_cast<float4*>(it.a + j);
      const float4 g = *reinterpret_cast<const float4*>(it.b + j);
      float4 m = *reinterpret_cast<float4*>(it.c + j);
      float4 v = *reinterpret_cast<float4*>(it.d + j);
      adamw_one(p.x, g.x, m.x, v.x, h);
      adamw_one(p.y, g.y, m.y, v.y, h);
      adamw_one(p.z, g.z, m.z, v.z, h);
      adamw_one(p.w, g.w, m.w, v.w, h);
      *reinterpret_cast<float4*>(it.a + j) = p;
      *reinterpret_cast<float4*>(it.c + j) = m;
      *reinterpret_cast<float4*>(it.d + j) = v;
    }
  } else {
    for (long long k = i; k < it.n && k < i + 8; ++k) adamw_one(it.a[k], it.b[k], it.c[k], it.d[k], h);
  }
}

__global__ void __launch_bounds__(256) ema_kernel(const OptItem* __restrict__ items, int n_items, float decay,
                                                  float w) {
  pdl_prologue();
  const OptItem it = find_item(items, n_items, blockIdx.x);
  const long long i = (long long)(blockIdx.x - it.first_block) * kOptChunk + (long long)threadIdx.x * 8;
  const uintptr_t al = reinterpret_cast<uintptr_t>(it.a) | reinterpret_cast<uintptr_t>(it.b);
  if (i + 8 <= it.n && (al & 15) == 0) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const long long j = i + half * 4;
      float4 e = *reinterpret_cast<float4*>(it.a + j);
      const float4 x = *reinterpret_cast<const float4*>(it.b + j);
      e.x = e.x * decay + w * x.x; e.y = e.y * decay + w * x.y; e.z = e.z * decay + w * x.z; e.w = e.w * decay + w * x.w;
      *reinterpret_cast<float4*>(it.a + j) = e;
    }
  } else {
    for (long long k = i; k < it.n && k < i + 8; ++k) it.a[k] = it.a[k] * decay + w * it.b[k];
  }
}

}  // namespace toist

using namespace toist;

extern "C" {

size_t toist_sizeof_opt_item(void) { return sizeof(OptItem); }

int toist_grad_sqnorm(const void* items_dev, int32_t n_items, int32_t total_blocks, float* partial, float max_norm,
                      float* norm_and_coef, void* stream) {
  TOIST_REQUIRE(items_dev && partial && norm_and_coef && n_items > 0 && total_blocks > 0, "toist_grad_sqnorm: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  TOIST_CHECK_CUDA(launch_pdl(grad_sqnorm_kernel, dim3(total_blocks), dim3(256), 0, st,
                              reinterpret_cast<const OptItem*>(items_dev), n_items, partial));
  TOIST_CHECK_CUDA(launch_pdl(grad_norm_finish_kernel, dim3(1), dim3(1024), 0, st, (const float*)partial, total_blocks,
                              max_norm, norm_and_coef));
  return TOIST_OK;
}

int toist_grad_clip_scale(const void* items_dev, int32_t n_items, int32_t total_blocks, const float* norm_and_coef,
                          void* stream) {
  TOIST_REQUIRE(items_dev && norm_and_coef && n_items > 0 && total_blocks > 0, "toist_grad_clip_scale: bad arguments");
  TOIST_CHECK_CUDA(launch_pdl(grad_scale_kernel, dim3(total_blocks), dim3(256), 0, (cudaStream_t)stream,
                              reinterpret_cast<const OptItem*>(items_dev), n_items, norm_and_coef));
  return TOIST_OK;
}

// hyper: n_groups x 8 floats {1 - lr * wd, 1 - beta1, beta2, 1 - beta2, eps, lr / (1 - beta1^t), sqrt(1 - beta2^t), 0}
int toist_adamw_step(const void* items_dev, int32_t n_items, int32_t total_blocks, const float* hyper_host,
                     int32_t n_groups, void* stream) {
  TOIST_REQUIRE(items_dev && hyper_host && n_items > 0 && total_blocks > 0, "toist_adamw_step: bad arguments");
  TOIST_REQUIRE(n_groups >= 1 && n_groups <= kMaxAdamRows, "toist_adamw_step: 1..%d hyper-parameter rows (got %d)", kMaxAdamRows,
                n_groups);
  AdamGroups g;
  memset(&g, 0, sizeof(g));
  memcpy(g.g, hyper_host, sizeof(AdamHyper) * n_groups);
  TOIST_CHECK_CUDA(launch_pdl(adamw_kernel, dim3(total_blocks), dim3(256), 0, (cudaStream_t)stream,
                              reinterpret_cast<const OptItem*>(items_dev), n_items, g));
  return TOIST_OK;
}

int toist_ema_update(const void* items_dev, int32_t n_items, int32_t total_blocks, float decay, float one_minus_decay,
                     void* stream) {
  TOIST_REQUIRE(items_dev && n_items > 0 && total_blocks > 0, "toist_ema_update: bad arguments");
  TOIST_CHECK_CUDA(launch_pdl(ema_kernel, dim3(total_blocks), dim3(256), 0, (cudaStream_t)stream,
                              reinterpret_cast<const OptItem*>(items_dev), n_items, decay, one_minus_decay));
  return TOIST_OK;
}

}  // extern "C"
