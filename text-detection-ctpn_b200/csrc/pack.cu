// Weight packing for the tensor-core layers: TF-layout float32 weights [taps][Cin][Cout] (HWIO flattened,
// lib/networks/network.py:160-170) -> bf16 planes [P][Cout_pad][taps][Cin] (K-major rows of the UMMA B operand).
#include "common.cuh"

namespace ctpn {

// ---- weight packing ---------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float *__restrict__ w, int taps, int cin, int cout, int cout_pad,
                                    int planes, __nv_bfloat16 *__restrict__ out) {
  const long long n = (long long)cout_pad * taps * cin;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ci = (int)(i % cin);
  const int tap = (int)((i / cin) % taps);
  const int co = (int)(i / ((long long)cin * taps));
  float v = co < cout ? w[((long long)tap * cin + ci) * cout + co] : 0.f;
  __nv_bfloat16 pl[3];
  split_planes(v, planes, pl);
  for (int p = 0; p < planes; ++p) out[(long long)p * n + i] = pl[p];
}

}  // namespace ctpn

using namespace ctpn;

extern "C" int ctpn_pack_weights(const float *w_tf, int taps, int cin, int cout, int cout_pad, int planes,
                                 void *w_planes_out, void *stream) {
  CTPN_REQUIRE(w_tf && w_planes_out, "ctpn_pack_weights: null pointer");
  CTPN_REQUIRE(taps > 0 && cin > 0 && cout > 0 && cout_pad >= cout, "ctpn_pack_weights: bad shape");
  CTPN_REQUIRE(planes >= 1 && planes <= 3, "ctpn_pack_weights: planes must be 1..3");
  const long long n = (long long)cout_pad * taps * cin;
  pack_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w_tf, taps, cin, cout, cout_pad, planes, reinterpret_cast<__nv_bfloat16 *>(w_planes_out));
  CTPN_LAUNCH_CHECK();
  return CTPN_OK;
}

