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

}  // namespace ctpn
