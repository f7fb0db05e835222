// Host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "toist_b200.h"

namespace toist {

// Thread-local error text returned by toist_last_error().
int set_error(int code, const char* fmt, ...);

#define TOIST_CHECK_CUDA(expr)                                                                        \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return ::toist::set_error(TOIST_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                __FILE__, __LINE__);                                                  \
  } while (0)

#define TOIST_REQUIRE(cond, ...)                                        \
  do {                                                                  \
    if (!(cond)) return ::toist::set_error(TOIST_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Encodes a rank-4 bf16 tensor map (SWIZZLE_128B, zero OOB fill).  dims/strides in elements (stride[0] == 1),
// box extents in elements *before* applying elem_stride (TMA loads ceil(box / elem_stride) per dimension).
int encode_tmap_bf16_4d(CUtensorMap* out, const void* ptr, const int64_t dim[4], const int64_t stride[4],
                        const uint32_t box[4], const uint32_t elem_stride[4]);

// Same for fp32 tensors (inner box <= 32 elements = one 128-byte swizzle row); used by the reduce-add epilogue.
int encode_tmap_f32_4d(CUtensorMap* out, const void* ptr, const int64_t dim[4], const int64_t stride[4],
                       const uint32_t box[4], const uint32_t elem_stride[4]);

// Kernel launch with programmatic dependent launch enabled (TOIST_PDL=0 in the environment restores plain stream
// serialisation for A/B measurements).  Every kernel launched through here calls pdl_wait() before touching memory.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// measurement hook (toist_debug_skip_gemm): when set, toist_gemm validates nothing and launches nothing
bool skip_gemm();
// measurement hook (toist_debug_gemm_trace): device buffer of 8 clock stamps per CTA, or nullptr
long long* gemm_trace_buffer();

}  // namespace toist
