// Box arithmetic shared by the matching-cost kernels (matcher.cu, distill.cu).
// Reference: util/box_ops.py:11-61 (box_cxcywh_to_xyxy, box_iou, generalized_box_iou).
#pragma once
#include "common.cuh"

namespace toist {

// fp32, round-to-nearest, no FMA contraction: the cost must follow the reference's operation order
// (matcher.py:71-81, box_ops.py:24-61) so that near-tie assignments agree.
struct BoxXYXY {
  float x0, y0, x1, y1;
};
__device__ __forceinline__ BoxXYXY to_xyxy(float cx, float cy, float w, float h) {
  BoxXYXY b;
  b.x0 = __fsub_rn(cx, __fmul_rn(0.5f, w));
  b.y0 = __fsub_rn(cy, __fmul_rn(0.5f, h));
  b.x1 = __fadd_rn(cx, __fmul_rn(0.5f, w));
  b.y1 = __fadd_rn(cy, __fmul_rn(0.5f, h));
  return b;
}
__device__ __forceinline__ float giou_xyxy(const BoxXYXY& a, const BoxXYXY& b) {
  const float area_a = __fmul_rn(__fsub_rn(a.x1, a.x0), __fsub_rn(a.y1, a.y0));
  const float area_b = __fmul_rn(__fsub_rn(b.x1, b.x0), __fsub_rn(b.y1, b.y0));
  const float iw = fmaxf(__fsub_rn(fminf(a.x1, b.x1), fmaxf(a.x0, b.x0)), 0.f);
  const float ih = fmaxf(__fsub_rn(fminf(a.y1, b.y1), fmaxf(a.y0, b.y0)), 0.f);
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  const float iou = __fdiv_rn(inter, uni);
  const float hw = fmaxf(__fsub_rn(fmaxf(a.x1, b.x1), fminf(a.x0, b.x0)), 0.f);
  const float hh = fmaxf(__fsub_rn(fmaxf(a.y1, b.y1), fminf(a.y0, b.y0)), 0.f);
  const float hull = __fmul_rn(hw, hh);
  return __fsub_rn(iou, __fdiv_rn(__fsub_rn(hull, uni), hull));
}

// torch.cdist(a, b, p=1) of two cxcywh boxes, accumulated in coordinate order
__device__ __forceinline__ float box_l1(const float* a, const float* b) {
  return __fadd_rn(__fadd_rn(__fadd_rn(fabsf(__fsub_rn(a[0], b[0])), fabsf(__fsub_rn(a[1], b[1]))),
                             fabsf(__fsub_rn(a[2], b[2]))),
                   fabsf(__fsub_rn(a[3], b[3])));
}

}  // namespace toist
