// Host-side helpers for TMA tensor maps.  cuTensorMapEncodeTiled is a driver-API function; it is
// resolved at run time through cudaGetDriverEntryPoint so that libctpn_b200.so has no link-time
// dependency on libcuda.so (the library must load on a machine without a GPU driver).
#pragma once
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace ctpn {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int tma_get_encode(EncodeTiledFn *out) {
  static EncodeTiledFn fn = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CTPN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
    if (!sym || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return CTPN_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  *out = fn;
  return CTPN_OK;
}

// bf16 tensor, 128-byte swizzle, zero fill for out-of-bounds elements
inline int tma_encode_bf16(EncodeTiledFn fn, CUtensorMap *m, void *addr, int rank, const cuuint64_t *dims,
                           const cuuint64_t *strides, const cuuint32_t *box) {
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, addr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return CTPN_ERR_CUDA;
  }
  return CTPN_OK;
}

}  // namespace ctpn
