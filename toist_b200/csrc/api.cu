// C-ABI plumbing: version, error text, device check, TMA tensor-map encoding through the driver entry point.
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "host_util.h"

namespace toist {

static thread_local char g_err[512] = "";
static int g_skip_gemm = 0;

bool skip_gemm() { return g_skip_gemm != 0; }
static long long* g_gemm_trace = nullptr;
long long* gemm_trace_buffer() { return g_gemm_trace; }

bool pdl_enabled() {
  static const bool on = []() {
    const char* e = getenv("TOIST_PDL");
    return !(e && atoi(e) == 0);
  }();
  return on;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // libcuda is resolved at run time through the runtime, so the .so links (and loads) on a CPU-only box.
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode_tmap_4d(CUtensorMap* out, const void* ptr, const int64_t dim[4], const int64_t stride[4],
                          const uint32_t box[4], const uint32_t elem_stride[4], int elem_bytes,
                          CUtensorMapDataType dtype);

int encode_tmap_bf16_4d(CUtensorMap* out, const void* ptr, const int64_t dim[4], const int64_t stride[4],
                        const uint32_t box[4], const uint32_t elem_stride[4]) {
  return encode_tmap_4d(out, ptr, dim, stride, box, elem_stride, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

int encode_tmap_f32_4d(CUtensorMap* out, const void* ptr, const int64_t dim[4], const int64_t stride[4],
                       const uint32_t box[4], const uint32_t elem_stride[4]) {
  return encode_tmap_4d(out, ptr, dim, stride, box, elem_stride, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

static int encode_tmap_4d(CUtensorMap* out, const void* ptr, const int64_t dim[4], const int64_t stride[4],
                          const uint32_t box[4], const uint32_t elem_stride[4], int elem_bytes,
                          CUtensorMapDataType dtype) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(TOIST_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  TOIST_REQUIRE(ptr != nullptr, "tensor map: null base pointer");
  TOIST_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tensor map: base %p not 16-byte aligned", ptr);
  TOIST_REQUIRE(stride[0] == 1, "tensor map: innermost stride must be 1 (got %lld)", (long long)stride[0]);
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < 4; ++i) {
    TOIST_REQUIRE(dim[i] >= 1, "tensor map: dim[%d]=%lld must be >= 1", i, (long long)dim[i]);
    gdim[i] = (cuuint64_t)dim[i];
    TOIST_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map: box[%d]=%u out of [1,256]", i, box[i]);
    bx[i] = box[i];
    es[i] = elem_stride[i];
    if (i > 0) {
      TOIST_REQUIRE((stride[i] * elem_bytes) % 16 == 0, "tensor map: stride[%d]=%lld elements is not a multiple of 16 bytes",
                    i, (long long)stride[i]);
      // a size-1 dimension may carry stride 0 in a torch view; TMA needs a positive multiple of 16 bytes
      int64_t s = stride[i] > 0 ? stride[i] : 16 / elem_bytes;
      gstr[i - 1] = (cuuint64_t)s * (cuuint64_t)elem_bytes;
    }
  }
  CUresult r = fn(out, dtype, 4, const_cast<void*>(ptr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TOIST_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): dims=[%lld,%lld,%lld,%lld] strides=[%lld,%lld,%lld,%lld] "
                     "box=[%u,%u,%u,%u] estr=[%u,%u,%u,%u]",
                     (int)r, (long long)dim[0], (long long)dim[1], (long long)dim[2], (long long)dim[3],
                     (long long)stride[0], (long long)stride[1], (long long)stride[2], (long long)stride[3], box[0],
                     box[1], box[2], box[3], es[0], es[1], es[2], es[3]);
  return TOIST_OK;
}

}  // namespace toist

extern "C" {

int toist_abi_version(void) { return TOIST_ABI_VERSION; }

const char* toist_last_error(void) { return toist::g_err; }

size_t toist_sizeof_gemm_desc(void) { return sizeof(toist_gemm_desc); }

int toist_debug_skip_gemm(int on) {
  const int prev = toist::g_skip_gemm;
  toist::g_skip_gemm = on;
  return prev;
}

int toist_debug_gemm_trace(void* buf) {
  toist::g_gemm_trace = reinterpret_cast<long long*>(buf);
  return TOIST_OK;
}

int toist_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

}  // extern "C"
