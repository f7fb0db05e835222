/*
 * toist_b200 — C ABI of the B200-native (sm_100a) kernels behind the TOIST / MDETR training hot path.
 *
 * The reference (AIR-DISCOVER/TOIST) is pure Python/PyTorch and has no FFI of its own; every entry point below
 * replaces a *library call site* inside the reference's hot path (cited per function as reference file:line).
 * The reference-side binding a maintainer would add is the ctypes stub in INTEGRATION.md (toist_b200/_lib.py is
 * the shipped copy of it).
 *
 * Conventions
 *   - plain C, POD structs of raw device pointers + sizes; no torch / C++ types cross this boundary
 *   - the caller allocates every output and workspace; kernels never allocate and never synchronise
 *   - every launch is asynchronous on the cudaStream_t passed as `stream` (a void* here)
 *   - return value: 0 on success, negative toist_status otherwise; toist_last_error() gives the thread-local text
 *   - bf16 tensors are raw uint16 storage (torch.bfloat16); "f32" is IEEE binary32
 */
#ifndef TOIST_B200_H_
#define TOIST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOIST_ABI_VERSION 1

typedef enum toist_status {
  TOIST_OK = 0,
  TOIST_ERR_INVALID = -1, /* bad argument (shape, alignment, null pointer) */
  TOIST_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed */
  TOIST_ERR_UNSUPPORTED = -3,
  TOIST_ERR_NUMERIC = -4 /* NaN / infeasible cost matrix (mirrors scipy's ValueError) */
} toist_status;

int toist_abi_version(void);
const char* toist_last_error(void);
/* sizeof(toist_gemm_desc) as compiled, so a foreign-language binding can verify its struct layout */
size_t toist_sizeof_gemm_desc(void);
/* 1 when the current device is compute capability 10.x (tcgen05 / TMEM present) */
int toist_device_ok(void);

/* ------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM engine (tcgen05.mma + TMEM accumulators, operands staged by TMA, SWIZZLE_128B).
 *
 * One kernel family serves: nn.Linear fwd/dgrad/wgrad (reference models/transformer.py:275-277,301,342-344,405,
 * 483; models/mdetr.py:347-357,1031; transformers RobertaModel), nn.MultiheadAttention projections and the
 * batched QK^T / PV products (transformer.py:273,298,337-338,378,394), the backbone / input_proj / mask-head
 * convolutions with FrozenBatchNorm + ReLU + residual folded into the epilogue (models/backbone.py:48-58,75;
 * models/mdetr.py:351; models/segmentation.py:178-195), and their data / weight gradients.
 *
 * Both operands are bf16 tensors described as 4-D views (dim[0] innermost, stride[0] == 1, all other strides
 * multiples of 8 elements, base 16-byte aligned).  The M side of the product is a *pixel tile*
 * (tile_x * tile_y * tile_n == 128 rows) so that a 3x3 tap is just a shifted TMA box with hardware zero fill.
 *
 *   mode FWD   : D[pix, n] = sum_tap sum_c A[c, x*sx+dx, y*sy+dy, n_img+dn] * B[tap.col + c, n, (y), (n_img)]
 *                A K-major (channels innermost), B K-major.
 *   mode DGRAD : same A traversal, B is MN-major: B[tap.col + n, c, (y), (n_img)]   (reduction index c on dim 1)
 *   mode WGRAD : D[m, tap.col + n] (+)= sum_pix A[m, pix] * B[n, pix*s + tap]  (both operands MN-major,
 *                reduction over pixel tiles of 64, optional split over grid.z with fp32 atomics)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct toist_tensor4 {
  const void* ptr;
  int64_t dim[4];    /* extents in elements, dim[0] innermost */
  int64_t stride[4]; /* strides in elements, stride[0] must be 1 */
} toist_tensor4;

typedef struct toist_tap {
  int16_t dx, dy, dn; /* coordinate offsets added to the pixel-tiled operand */
  int16_t pad_;
  int32_t col; /* FWD/DGRAD: column offset into B dim0; WGRAD: column offset into the output */
} toist_tap;

enum { TOIST_GEMM_FWD = 0, TOIST_GEMM_DGRAD = 1, TOIST_GEMM_WGRAD = 2 };
enum { TOIST_ACT_NONE = 0, TOIST_ACT_RELU = 1, TOIST_ACT_GELU = 2, TOIST_ACT_SIGMOID = 3 };
enum { TOIST_BF16 = 0, TOIST_F32 = 1 };
#define TOIST_MAX_TAPS 49

typedef struct toist_gemm_desc {
  int32_t mode;
  toist_tensor4 a, b;
  /* pixel space of the M side (FWD/DGRAD) or of the reduction (WGRAD) */
  int32_t ext_x, ext_y, ext_n;    /* logical extents iterated by tiles */
  int32_t tile_x, tile_y, tile_n; /* product 128 (FWD/DGRAD) or 64 (WGRAD) */
  int32_t stride_x, stride_y;     /* element stride of the pixel-tiled operand (conv stride), 1 or 2 */
  int32_t n_cols;                 /* N extent of the product (FWD/DGRAD: output columns; WGRAD: B rows used) */
  int32_t m_rows;                 /* WGRAD only: M extent (rows of D) */
  int32_t k_per_tap;              /* FWD/DGRAD: reduction length per tap (channels) */
  int32_t n_taps;
  toist_tap taps[TOIST_MAX_TAPS];
  int32_t b_batched; /* FWD/DGRAD: B dims 2,3 indexed by the tile's (y, n_img); WGRAD: unused */
  int32_t batch_y, batch_n; /* WGRAD: extra batch grid; A/B dims 2,3 and output offset by (by, bn) */
  int32_t splits;    /* WGRAD: split the pixel reduction over this many CTAs (atomic fp32 accumulate) */
  /* epilogue: v = acc * alpha; v = v * col_scale[n] + col_shift[n]; v *= row_scale[m]; v += res; v *= (mask > 0);
   *           aux = v (pre-activation, bf16); v = act(v); out (=|+=) v                                          */
  void* out;
  int32_t out_dtype; /* TOIST_BF16 | TOIST_F32 */
  int64_t out_sx, out_sy, out_sn; /* element strides of the output along pixel x / y / n (WGRAD: sx = row stride,
                                     sy / sn = batch strides); columns are contiguous */
  float alpha;
  const float* col_scale;
  const float* col_shift;
  const float* row_scale;
  const void* res; /* same addressing as out */
  int32_t res_dtype;
  const void* mask; /* bf16, same addressing as out */
  void* aux;        /* bf16, same addressing as out */
  int32_t act;
  int32_t accumulate; /* 1: out += v using fp32 atomics (out must be f32) */
} toist_gemm_desc;

/* Launches the product described by `d` on `stream`. */
int toist_gemm(const toist_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOIST_B200_H_ */
