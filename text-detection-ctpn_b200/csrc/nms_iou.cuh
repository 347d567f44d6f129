// Exact float32 IoU shared by the generic NMS (nms.cu) and the column-wise NMS (proposal.cu).
// Same operation order as lib/utils/nms_kernel.cu:24-32 / lib/fast_rcnn/nms_wrapper.py:30,37-44,
// no FMA contraction, IEEE division.
#pragma once
#include <cuda_runtime.h>

namespace ctpn {

__device__ __forceinline__ float box_area(float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

__device__ __forceinline__ float iou_exact(float4 a, float sa, float4 b, float sb) {
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
  float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
  float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
}

// The decision `iou_exact(a, b) > thresh` without the IEEE division for all but the borderline pairs: disjoint boxes have
// inter = 0 (never above a non-negative threshold), and for the rest the sign of inter - thresh * union decides unless it lies
// within 2^-20 of inter, where the correctly rounded quotient is evaluated exactly as before.  Bit-identical decisions
// (tests/test_nms_gpu.py, tests/test_proposals_gpu.py compare keep lists with the CPU oracle).
__device__ __forceinline__ bool iou_above(float4 a, float sa, float4 b, float sb, float thresh) {
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
  float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
  float inter = __fmul_rn(w, h);
  if (thresh >= 0.f && inter == 0.f) return false;          // 0 / u, NaN and -0 are all "not above"
  float u = __fsub_rn(__fadd_rn(sa, sb), inter);
  if (u > 0.f && thresh >= 0.f) {
    const float d = fmaf(-thresh, u, inter);                 // ~ inter - thresh * u (errors ~2^-23 of inter + thresh * u)
    const float margin = 9.5367431640625e-7f * fmaxf(inter, u);   // 2^-20
    if (d > margin) return true;
    if (d < -margin) return false;
  }
  return __fdiv_rn(inter, u) > thresh;
}

}  // namespace ctpn
