// The two ends of the hot path, HBM-bound byte / float work, one pass each:
//   * batch assembly: per-image uint8 HWC (or fp32 CHW) tensors -> ToTensor + Normalize + zero padding + padding mask
//     in one launch (reference datasets/transforms.py:257-272 on the CPU workers, util/misc.py:185-209 one copy + one
//     mask fill per image)
//   * PostProcess: softmax "not-background" score + cxcywh -> absolute xyxy boxes (models/postprocessors.py:15-58)
//   * PostProcessSegm: the reference's TWO chained bilinear interpolations (mask -> padded batch size, crop, -> original
//     image size), sigmoid and threshold fused into one read of the low-resolution mask (models/postprocessors.py:61-109)
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace toist {

// ------------------------------------------------------------------------------------------------ batch assembly
// out [B, 3, H, W] f32, mask [B, H, W] u8 (1 = padding).  One thread per output pixel (all three channels).
__global__ void pad_normalize_u8_kernel(const unsigned long long* __restrict__ ptrs, const int* __restrict__ hw,
                                        float* __restrict__ out, uint8_t* __restrict__ mask, int B, int H, int W,
                                        float m0, float m1, float m2, float s0, float s1, float s2) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long plane = (long long)H * W;
  if (i >= (long long)B * plane) return;
  const int b = (int)(i / plane);
  const int r = (int)(i % plane);
  const int y = r / W, x = r % W;
  const int h = hw[2 * b], w = hw[2 * b + 1];
  float* o = out + (long long)b * 3 * plane + r;
  if (y < h && x < w) {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(ptrs[b]) + ((long long)y * w + x) * 3;
    // ToTensor: x / 255 (true division), Normalize: (x - mean) / std, each rounded once like torch's tensor ops
    o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[0], 255.f), m0), s0);
    o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[1], 255.f), m1), s1);
    o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[2], 255.f), m2), s2);
    mask[i] = 0;
  } else {
    o[0] = 0.f;
    o[plane] = 0.f;
    o[2 * plane] = 0.f;
    mask[i] = 1;
  }
}

// images already fp32 [C, h_b, w_b] (normalised by the data pipeline): zero padding + mask only (util/misc.py:185-209)
__global__ void pad_f32_kernel(const unsigned long long* __restrict__ ptrs, const int* __restrict__ hw,
                               float* __restrict__ out, uint8_t* __restrict__ mask, int B, int Cc, int H, int W) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long plane = (long long)H * W;
  if (i >= (long long)B * plane) return;
  const int b = (int)(i / plane);
  const int r = (int)(i % plane);
  const int y = r / W, x = r % W;
  const int h = hw[2 * b], w = hw[2 * b + 1];
  const bool in = y < h && x < w;
  const float* src = reinterpret_cast<const float*>(ptrs[b]);
  float* o = out + (long long)b * Cc * plane + r;
  for (int c = 0; c < Cc; ++c) o[c * plane] = in ? src[((long long)c * h + y) * w + x] : 0.f;
  mask[i] = in ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ PostProcess
// One warp per (image, query): scores = 1 - softmax(logits)[-1]; boxes cxcywh -> xyxy scaled to (w, h, w, h).
__global__ void postprocess_boxes_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                         const float* __restrict__ sizes_f, const long long* __restrict__ sizes_i,
                                         const float* __restrict__ is_final, float* __restrict__ scores,
                                         long long* __restrict__ labels, float* __restrict__ out_boxes,
                                         float* __restrict__ scores_refexp, int B, int Q, int Cc) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * Q) return;
  const float* l = logits + (long long)row * Cc;
  float mx = -INFINITY;
  for (int c = lane; c < Cc; c += 32) mx = fmaxf(mx, l[c]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
  for (int c = lane; c < Cc; c += 32) se += expf(l[c] - mx);
  se = warp_sum(se);
  if (lane == 0) {
    const float p_last = expf(l[Cc - 1] - mx) / se;
    const float sc = 1.f - p_last;
    scores[row] = sc;
    labels[row] = 1;
    if (scores_refexp != nullptr) scores_refexp[row] = sc * (1.f / (1.f + expf(-is_final[row])));
    const int b = row / Q;
    float ih, iw;
    if (sizes_i != nullptr) {
      ih = (float)sizes_i[2 * b];
      iw = (float)sizes_i[2 * b + 1];
    } else {
      ih = sizes_f[2 * b];
      iw = sizes_f[2 * b + 1];
    }
    const float cx = boxes[row * 4 + 0], cy = boxes[row * 4 + 1], w = boxes[row * 4 + 2], h = boxes[row * 4 + 3];
    // util/box_ops.py:11-14, no fused multiply-add: x_c - 0.5 * w etc., then * scale
    out_boxes[row * 4 + 0] = __fmul_rn(__fsub_rn(cx, __fmul_rn(0.5f, w)), iw);
    out_boxes[row * 4 + 1] = __fmul_rn(__fsub_rn(cy, __fmul_rn(0.5f, h)), ih);
    out_boxes[row * 4 + 2] = __fmul_rn(__fadd_rn(cx, __fmul_rn(0.5f, w)), iw);
    out_boxes[row * 4 + 3] = __fmul_rn(__fadd_rn(cy, __fmul_rn(0.5f, h)), ih);
  }
}

// ------------------------------------------------------------------------------------------------ PostProcessSegm
// torch's upsample_bilinear2d source index (align_corners = false, size given): src = (dst + .5) * in / out - .5, >= 0
__device__ __forceinline__ void bil_src(int d, int in, int out, int& i0, int& i1, float& l0, float& l1) {
  float s = __fmaf_rn((float)d + 0.5f, (float)in / (float)out, -0.5f);
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i0 = i0 < in - 1 ? i0 : in - 1;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

// value of the FIRST interpolation (low-res mask (hm, wm) -> (H1, W1)) at integer position (y, x)
__device__ __forceinline__ float stage1(const float* __restrict__ pm, int hm, int wm, int H1, int W1, int y, int x) {
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bil_src(y, hm, H1, y0, y1, ly0, ly1);
  bil_src(x, wm, W1, x0, x1, lx0, lx1);
  return ly0 * (lx0 * pm[y0 * wm + x0] + lx1 * pm[y0 * wm + x1]) + ly1 * (lx0 * pm[y1 * wm + x0] + lx1 * pm[y1 * wm + x1]);
}

// pred [Q, hm, wm] f32 of ONE image -> out [Q, OH, OW] u8 = sigmoid(interp2(crop(interp1(pred)))) > threshold
__global__ void postprocess_masks_kernel(const float* __restrict__ pred, uint8_t* __restrict__ out, int Q, int hm, int wm,
                                         int H1, int W1, int ch, int cw, int OH, int OW, float threshold) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long plane = (long long)OH * OW;
  if (i >= (long long)Q * plane) return;
  const int q = (int)(i / plane);
  const int r = (int)(i % plane);
  const int y = r / OW, x = r % OW;
  const float* pm = pred + (long long)q * hm * wm;
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bil_src(y, ch, OH, y0, y1, ly0, ly1);  // second interpolation: the (ch, cw) crop of the stage-1 map -> (OH, OW)
  bil_src(x, cw, OW, x0, x1, lx0, lx1);
  const float v00 = stage1(pm, hm, wm, H1, W1, y0, x0), v01 = stage1(pm, hm, wm, H1, W1, y0, x1);
  const float v10 = stage1(pm, hm, wm, H1, W1, y1, x0), v11 = stage1(pm, hm, wm, H1, W1, y1, x1);
  const float v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
  const float p = 1.f / (1.f + expf(-v));
  out[i] = p > threshold ? 1 : 0;
}

}  // namespace toist

using namespace toist;

extern "C" {

int toist_pad_normalize_u8(const uint64_t* image_ptrs, const int32_t* image_hw, float* out, uint8_t* mask, int32_t batch,
                           int32_t height, int32_t width, const float* mean3_host, const float* std3_host, void* stream) {
  TOIST_REQUIRE(image_ptrs && image_hw && out && mask && mean3_host && std3_host, "toist_pad_normalize_u8: null pointer");
  TOIST_REQUIRE(batch >= 1 && height >= 1 && width >= 1, "toist_pad_normalize_u8: empty batch");
  const long long total = (long long)batch * height * width;
  TOIST_CHECK_CUDA(launch_pdl(pad_normalize_u8_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                              (cudaStream_t)stream, (const unsigned long long*)image_ptrs, (const int*)image_hw, out, mask,
                              (int)batch, (int)height, (int)width, mean3_host[0], mean3_host[1], mean3_host[2],
                              std3_host[0], std3_host[1], std3_host[2]));
  return TOIST_OK;
}

int toist_pad_batch_f32(const uint64_t* image_ptrs, const int32_t* image_hw, float* out, uint8_t* mask, int32_t batch,
                        int32_t channels, int32_t height, int32_t width, void* stream) {
  TOIST_REQUIRE(image_ptrs && image_hw && out && mask, "toist_pad_batch_f32: null pointer");
  TOIST_REQUIRE(batch >= 1 && channels >= 1 && height >= 1 && width >= 1, "toist_pad_batch_f32: empty batch");
  const long long total = (long long)batch * height * width;
  TOIST_CHECK_CUDA(launch_pdl(pad_f32_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
                              (const unsigned long long*)image_ptrs, (const int*)image_hw, out, mask, (int)batch,
                              (int)channels, (int)height, (int)width));
  return TOIST_OK;
}

int toist_postprocess_boxes(const float* logits, const float* boxes, const float* sizes_f32, const int64_t* sizes_i64,
                            const float* is_final, float* scores, int64_t* labels, float* out_boxes, float* scores_refexp,
                            int32_t batch, int32_t n_queries, int32_t n_classes, void* stream) {
  TOIST_REQUIRE(logits && boxes && scores && labels && out_boxes, "toist_postprocess_boxes: null pointer");
  TOIST_REQUIRE((sizes_f32 != nullptr) != (sizes_i64 != nullptr), "toist_postprocess_boxes: pass exactly one size array");
  TOIST_REQUIRE((is_final == nullptr) == (scores_refexp == nullptr), "toist_postprocess_boxes: is_final needs scores_refexp");
  const int rows = batch * n_queries;
  if (rows == 0) return TOIST_OK;
  TOIST_CHECK_CUDA(launch_pdl(postprocess_boxes_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0,
                              (cudaStream_t)stream, logits, boxes, sizes_f32, (const long long*)sizes_i64, is_final,
                              scores, (long long*)labels, out_boxes, scores_refexp, (int)batch, (int)n_queries,
                              (int)n_classes));
  return TOIST_OK;
}

int toist_postprocess_masks(const float* pred_masks, uint8_t* out, int32_t n_queries, int32_t mask_h, int32_t mask_w,
                            int32_t stage1_h, int32_t stage1_w, int32_t crop_h, int32_t crop_w, int32_t out_h,
                            int32_t out_w, float threshold, void* stream) {
  TOIST_REQUIRE(pred_masks && out, "toist_postprocess_masks: null pointer");
  TOIST_REQUIRE(mask_h >= 1 && mask_w >= 1 && stage1_h >= 1 && stage1_w >= 1 && out_h >= 1 && out_w >= 1,
                "toist_postprocess_masks: empty size");
  TOIST_REQUIRE(crop_h >= 1 && crop_h <= stage1_h && crop_w >= 1 && crop_w <= stage1_w,
                "toist_postprocess_masks: crop (%d, %d) outside the stage-1 map (%d, %d)", crop_h, crop_w, stage1_h, stage1_w);
  const long long total = (long long)n_queries * out_h * out_w;
  if (total == 0) return TOIST_OK;
  TOIST_CHECK_CUDA(launch_pdl(postprocess_masks_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                              (cudaStream_t)stream, pred_masks, out, (int)n_queries, (int)mask_h, (int)mask_w,
                              (int)stage1_h, (int)stage1_w, (int)crop_h, (int)crop_w, (int)out_h, (int)out_w, threshold));
  return TOIST_OK;
}

}  // extern "C"
