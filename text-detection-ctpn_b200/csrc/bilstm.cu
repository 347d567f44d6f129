// BiLSTM recurrence over the feature-map width (lib/networks/network.py:93-101: LSTMCell(128)
// forward and backward over W for every feature-map row, zero initial state).
//
// The input projection x.Wx + b for both directions is a plain GEMM (done by ctpn_conv3x3 with
// taps = 1); this kernel runs the sequential part.  The recurrent matrix Wh is 128 x 512 float32
// = 256 KiB, more than one SM's shared memory, so a 2-CTA thread-block cluster splits the hidden
// units: CTA r keeps the 4 x 64 gate columns of units [64r, 64r+64) (128 KiB) resident in shared
// memory for the whole sequence and the two CTAs exchange their halves of h_t through
// distributed shared memory once per step.  Each cluster advances RG independent rows of one
// direction, so Wh is read from shared memory once per RG rows.  All arithmetic is float32; the mat-vec uses the
// packed fma.rn.f32x2 of sm_100 (two k per instruction: even-k and odd-k partial sums, added at the end).
//
// TF 1.3 LSTMCell: gates (i, j, f, o) = [x, h] . kernel + bias;
//   c = sigmoid(f + 1) * c + sigmoid(i) * tanh(j);  h = sigmoid(o) * tanh(c).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ctpn {

constexpr int kHid = 128, kGates = 512, kHalf = 64, kLocalCols = 256;

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int RG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
bilstm_kernel(const float *__restrict__ xproj, const float *__restrict__ wh_fw, const float *__restrict__ wh_bw,
              __nv_bfloat16 *__restrict__ out, int R, int W, int planes) {
  extern __shared__ __align__(16) float smem[];
  float *Ws = smem;                          // [128][256]  recurrent weights of this CTA's 64 units
  float *hbuf = Ws + kHid * kLocalCols;      // [2][RG][128] full hidden state, double buffered
  float *gates = hbuf + 2 * RG * kHid;       // [RG][256]    pre-activations of this CTA's columns
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x >> 1;                       // cluster index
  const int groups = (R + RG - 1) / RG;
  const int dir = cid / groups, row0 = (cid % groups) * RG;
  const int t = threadIdx.x;
  // mat-vec phase: thread = 4 adjacent local columns x RG/4 rows (register tile: 4 + RG/4 shared-memory loads
  // per 4 * RG/4 * 4 FMAs); cell phase: thread = unit ul, rows (t >> 6) + 4q
  constexpr int RT = RG / 4;
  const int cg = t & 63, rg = t >> 6;                    // column group (4 columns), row group (RT rows)
  const int lc0 = cg * 4;                                // first local column; gate = lc0 >> 6, unit = lc0 & 63
  const int gcol0 = (lc0 >> 6) * kHid + rank * kHalf + (lc0 & 63);   // column in the 512-wide gate vector
  const int ul = t & 63;
  const float *wh = dir ? wh_bw : wh_fw;
  // Shared-memory layout of the weights for packed (f32x2) FMAs: [column pair half][k pair][column group][4] with
  // the 4 floats = (col a: k even, k odd; col b: k even, k odd), so one LDS.128 hands a thread two (k, k+1) weight
  // pairs and consecutive lanes read consecutive 16-byte chunks (conflict-free).
  for (int i = t; i < kHid * kLocalCols; i += 256) {
    const int k = i >> 8, lc = i & 255;
    const int cgi = lc >> 2, cc = lc & 3;
    Ws[(((cc >> 1) * (kHid / 2) + (k >> 1)) * 64 + cgi) * 4 + (cc & 1) * 2 + (k & 1)] =
        wh[k * kGates + (lc >> 6) * kHid + rank * kHalf + (lc & 63)];
  }
  for (int i = t; i < 2 * RG * kHid; i += 256) hbuf[i] = 0.f;
  float *peer_h = cluster.map_shared_rank(hbuf, rank ^ 1);
  float c_state[RG / 4];
#pragma unroll
  for (int q = 0; q < RG / 4; ++q) c_state[q] = 0.f;
  cluster.sync();

  const long long plane_stride = (long long)R * W * 2 * kHid;
  float4 xnext[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int row = row0 + rg * RT + r;
    xnext[r] = row < R ? __ldg(reinterpret_cast<const float4 *>(xproj + ((long long)row * W + (dir ? W - 1 : 0)) * (2 * kGates) + dir * kGates + gcol0))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int step = 0; step < W; ++step) {
    const int tpos = dir ? W - 1 - step : step;
    const float *hc = hbuf + (step & 1) * RG * kHid;
    float *hn = hbuf + ((step + 1) & 1) * RG * kHid;
    float *hn_peer = peer_h + ((step + 1) & 1) * RG * kHid;
    // acc[r][c] = (sum over even k, sum over odd k) for local column lc0 + c: one FFMA2 (sm_100 fma.rn.f32x2)
    // advances two k at once with the natural register pairs (h[k], h[k+1]) x (w[k][c], w[k+1][c]).
    float2 acc[RT][4];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float4 x = xnext[r];
      acc[r][0] = make_float2(x.x, 0.f); acc[r][1] = make_float2(x.y, 0.f);
      acc[r][2] = make_float2(x.z, 0.f); acc[r][3] = make_float2(x.w, 0.f);
    }
    if (step + 1 < W) {   // x-projection of the next step: in flight during this step's mat-vec
      const int tn = dir ? W - 2 - step : step + 1;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const int row = row0 + rg * RT + r;
        if (row < R) xnext[r] = __ldg(reinterpret_cast<const float4 *>(xproj + ((long long)row * W + tn) * (2 * kGates) + dir * kGates + gcol0));
      }
    }
    const float4 *w_lo = reinterpret_cast<const float4 *>(Ws) + cg;                       // columns lc0, lc0 + 1
    const float4 *w_hi = reinterpret_cast<const float4 *>(Ws) + (kHid / 2) * 64 + cg;     // columns lc0 + 2, lc0 + 3
#pragma unroll 2
    for (int k = 0; k < kHid; k += 4) {
      const float4 wa0 = w_lo[(k >> 1) * 64], wa1 = w_lo[((k >> 1) + 1) * 64];
      const float4 wb0 = w_hi[(k >> 1) * 64], wb1 = w_hi[((k >> 1) + 1) * 64];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float4 h4 = *reinterpret_cast<const float4 *>(hc + (rg * RT + r) * kHid + k);
        const float2 h01 = make_float2(h4.x, h4.y), h23 = make_float2(h4.z, h4.w);
        acc[r][0] = __ffma2_rn(h01, make_float2(wa0.x, wa0.y), acc[r][0]);
        acc[r][1] = __ffma2_rn(h01, make_float2(wa0.z, wa0.w), acc[r][1]);
        acc[r][2] = __ffma2_rn(h01, make_float2(wb0.x, wb0.y), acc[r][2]);
        acc[r][3] = __ffma2_rn(h01, make_float2(wb0.z, wb0.w), acc[r][3]);
        acc[r][0] = __ffma2_rn(h23, make_float2(wa1.x, wa1.y), acc[r][0]);
        acc[r][1] = __ffma2_rn(h23, make_float2(wa1.z, wa1.w), acc[r][1]);
        acc[r][2] = __ffma2_rn(h23, make_float2(wb1.x, wb1.y), acc[r][2]);
        acc[r][3] = __ffma2_rn(h23, make_float2(wb1.z, wb1.w), acc[r][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r)
      *reinterpret_cast<float4 *>(gates + (rg * RT + r) * kLocalCols + lc0) =
          make_float4(acc[r][0].x + acc[r][0].y, acc[r][1].x + acc[r][1].y, acc[r][2].x + acc[r][2].y, acc[r][3].x + acc[r][3].y);
    __syncthreads();
    // cell update: thread -> unit ul, rows (t>>6) + 4q
#pragma unroll
    for (int q = 0; q < RG / 4; ++q) {
      const int r = (t >> 6) + 4 * q;
      const float gi = gates[r * kLocalCols + ul], gj = gates[r * kLocalCols + 64 + ul];
      const float gf = gates[r * kLocalCols + 128 + ul], go = gates[r * kLocalCols + 192 + ul];
      const float c = sigmoidf_acc(gf + 1.0f) * c_state[q] + sigmoidf_acc(gi) * tanhf(gj);
      const float h = sigmoidf_acc(go) * tanhf(c);
      c_state[q] = c;
      const int u = rank * kHalf + ul;
      hn[r * kHid + u] = h;
      hn_peer[r * kHid + u] = h;
      const int row = row0 + r;
      if (row < R) {
        __nv_bfloat16 pl[3];
        split_planes(h, planes, pl);
        const long long o = ((long long)row * W + tpos) * (2 * kHid) + dir * kHid + u;
        for (int p = 0; p < planes; ++p) out[p * plane_stride + o] = pl[p];
      }
    }
    cluster.sync();   // h_{t} of both halves visible in both CTAs; also orders the gates[] reuse
  }
}

template <int RG>
static int launch_bilstm(const float *xproj, const float *wh_fw, const float *wh_bw, void *out, int R, int W, int planes,
                         cudaStream_t st) {
  const size_t smem = (size_t)(kHid * kLocalCols + 2 * RG * kHid + RG * kLocalCols) * sizeof(float);
  CTPN_CUDA(cudaFuncSetAttribute(bilstm_kernel<RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int groups = (R + RG - 1) / RG;
  ProfScope prof("bilstm_recurrent", 2.0 * 2.0 * R * W * 128.0 * 512.0, st);
  bilstm_kernel<RG><<<2 * 2 * groups, 256, smem, st>>>(xproj, wh_fw, wh_bw, (__nv_bfloat16 *)out, R, W, planes);
  CTPN_LAUNCH_CHECK();
  return CTPN_OK;
}

}  // namespace ctpn

using namespace ctpn;

extern "C" int ctpn_bilstm_recurrent(const float *xproj, const float *wh_fw, const float *wh_bw, void *out_planes, int R,
                                     int W, int planes, void *stream) {
  CTPN_REQUIRE(xproj && wh_fw && wh_bw && out_planes, "ctpn_bilstm_recurrent: null pointer");
  CTPN_REQUIRE(R > 0 && W > 0, "ctpn_bilstm_recurrent: bad shape R=%d W=%d", R, W);
  CTPN_REQUIRE(planes >= 1 && planes <= 3, "ctpn_bilstm_recurrent: planes must be 1..3");
  cudaStream_t st = (cudaStream_t)stream;
  // rows per cluster: as many as keeps every SM busy in a single wave (148 SMs = 74 clusters per direction pair)
  if (R >= 32 * 37) return launch_bilstm<32>(xproj, wh_fw, wh_bw, out_planes, R, W, planes, st);
  if (R >= 16 * 37) return launch_bilstm<16>(xproj, wh_fw, wh_bw, out_planes, R, W, planes, st);
  if (R >= 8 * 37) return launch_bilstm<8>(xproj, wh_fw, wh_bw, out_planes, R, W, planes, st);
  return launch_bilstm<4>(xproj, wh_fw, wh_bw, out_planes, R, W, planes, st);
}
