// 3x3 SAME convolution / 1x1 matmul on bf16 "planes" with 5th-generation tensor cores.
//
// Implicit GEMM: M = output pixels (a TH x TW patch of one image per CTA tile, TH*TW = 128),
// N = output channels (BN per tile), K = taps * Cin in blocks of 64 channels.
//   * A operand: for each tap the TH x TW x 64 input patch shifted by (dy, dx) is fetched with
//     ONE 4-D TMA box load from the NHWC tensor; out-of-image elements are zero-filled by the
//     TMA unit, which implements SAME padding and ragged image edges without any halo buffer.
//     The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, i.e. exactly
//     the canonical K-major UMMA layout.
//   * B operand: weights [P][Cout][taps][Cin] viewed as a 2-D K-major matrix, 2-D TMA box.
//   * D: float32 accumulators in TMEM (double buffered: the epilogue of tile i overlaps the
//     MMAs of tile i+1), tcgen05.mma issued by one thread, completion via tcgen05.commit.
//   * float32-faithful mode: activations/weights are split into P bf16 planes; all plane pairs
//     (i, j) with i + j < P are accumulated into the same TMEM tile (1, 3 or 6 MMAs per k-step).
//   * epilogue (4 warps = 128 TMEM lanes): tcgen05.ld -> +bias -> ReLU -> optional fused 2x2
//     max-pool (warp shuffles: the 2x2 window of a pixel lives in lanes l, l^1, l^TW, l^TW^1)
//     -> re-split into planes -> 16-byte global stores (or float32 output).
// Persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer and
// TMEM owner, warps 2-5 epilogue.  Reference semantics: lib/networks/network.py:160-196.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ctpn {

struct ConvTcParams {
  int B, H, W, Cin, Cout, taps, planes, flags;
  int tiles_x, tiles_y, tiles_n, total_tiles;
  int TH, TW, tw_log2;
  int cout_pad;
  int Ho, Wo;
  int stages;
  const float *bias;
  void *out;
  long long out_plane_stride;   // elements between output planes
};

constexpr int kTcThreads = 192;
constexpr int kABytes = 128 * 128;   // 128 pixels x 64 bf16
constexpr int kMaxStages = 8;

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const ConvTcParams p) {
  using namespace ptx;
  constexpr int kBBytes = BN * 128;
  constexpr uint32_t kIdesc = umma_idesc_bf16(128, BN);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t stage0 = (raw + 1023u) & ~1023u;
  const int P = p.planes;
  const uint32_t stage_bytes = (uint32_t)P * (kABytes + kBBytes);
  uint8_t *ctrl = smem_raw + (stage0 - raw) + (size_t)p.stages * stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(ctrl);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kMaxStages;
  const uint32_t tfull0 = empty0 + 8 * kMaxStages, tempty0 = tfull0 + 16;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ctrl + 8 * (2 * kMaxStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);
    }
    fence_mbar_init();
  }
  // TMEM: two tile buffers (epilogue of tile i overlaps the MMAs of tile i+1).  With P > 1 every
  // buffer holds TWO accumulators: "main" receives the plane-0 x plane-0 products, "cross" all the
  // 2^-8 .. 2^-16 smaller cross-plane products.  The tensor core truncates (does not round) when it
  // adds into the float32 accumulator; keeping the small terms out of the big accumulator keeps
  // their truncation error proportional to THEIR magnitude (measured: 3-6x lower end-to-end error).
  const uint32_t acc_cols = (P > 1 ? 2u : 1u) * BN;
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 2 * acc_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblocks = p.Cin / 64;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.tiles_n, mt = tile / p.tiles_n;
        const int b = mt / tiles_per_img, r = mt % tiles_per_img;
        const int y0 = (r / p.tiles_x) * p.TH, x0 = (r % p.tiles_x) * p.TW;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            mbar_arrive_expect_tx(full0 + 8 * s, stage_bytes);
            const uint32_t sa = stage0 + (uint32_t)s * stage_bytes;
            const uint32_t sb = sa + (uint32_t)P * kABytes;
            for (int pl = 0; pl < P; ++pl)
              tma_load_4d(&tmap_a, full0 + 8 * s, sa + pl * kABytes, kb * 64, x0 + dx, y0 + dy, pl * p.B + b);
            for (int pl = 0; pl < P; ++pl)
              tma_load_2d(&tmap_b, full0 + 8 * s, sb + pl * kBBytes, tap * p.Cin + kb * 64, pl * p.cout_pad + nt * BN);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      int s = 0, a = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty0 + 8 * a, aph ^ 1u);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)a * acc_cols, d_cross = d_main + BN;
        uint32_t accum_main = 0, accum_cross = 0;
        const int ksteps = p.taps * kblocks;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = stage0 + (uint32_t)s * stage_bytes;
          const uint32_t sb = sa + (uint32_t)P * kABytes;
          for (int i = 0; i < P; ++i) {
            for (int j = 0; j < P - i; ++j) {
              const uint64_t da = umma_desc_k_sw128(sa + i * kABytes);
              const uint64_t db = umma_desc_k_sw128(sb + j * kBBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {   // 64 / UMMA_K(16); +32 B == +2 in the >>4 address field
                if (i + j == 0) { mma_bf16_ss(d_main, da + 2ull * k, db + 2ull * k, kIdesc, accum_main); accum_main = 1; }
                else { mma_bf16_ss(d_cross, da + 2ull * k, db + 2ull * k, kIdesc, accum_cross); accum_cross = 1; }
              }
            }
          }
          mma_commit(empty0 + 8 * s);   // frees the smem stage when these MMAs have read it
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        mma_commit(tfull0 + 8 * a);      // accumulator complete -> epilogue
        a ^= 1;
        if (a == 0) aph ^= 1u;
      }
    }
  } else {
    // ===== epilogue warps (TMEM lane quarter = warp id % 4) =====
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int th = m >> p.tw_log2, tw = m & (p.TW - 1);
    const bool pool = (p.flags & CTPN_F_POOL) != 0, relu = (p.flags & CTPN_F_RELU) != 0;
    const bool out_f32 = (p.flags & CTPN_F_OUT_F32) != 0;
    int a = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.tiles_n, mt = tile / p.tiles_n;
      const int b = mt / tiles_per_img, r = mt % tiles_per_img;
      const int y = (r / p.tiles_x) * p.TH + th, x = (r % p.tiles_x) * p.TW + tw;
      bool ok;
      int oy, ox;
      if (pool) {
        oy = y >> 1; ox = x >> 1;
        ok = !(th & 1) && !(tw & 1) && oy < p.Ho && ox < p.Wo;
      } else {
        oy = y; ox = x;
        ok = y < p.H && x < p.W;
      }
      const long long pix = ((long long)b * p.Ho + oy) * p.Wo + ox;
      mbar_wait(tfull0 + 8 * a, aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)a * acc_cols;
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t rr[32];
        tmem_ld_32x32(taddr + chunk * 32, rr);
        tmem_ld_wait();
        const int c0 = nt * BN + chunk * 32;
        float v[32];
        if (P > 1) {
          uint32_t rc[32];
          tmem_ld_32x32(taddr + BN + chunk * 32, rc);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]) + __uint_as_float(rc[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bq = __ldg(reinterpret_cast<const float4 *>(p.bias + c0) + q);
          v[4 * q + 0] += bq.x;
          v[4 * q + 1] += bq.y;
          v[4 * q + 2] += bq.z;
          v[4 * q + 3] += bq.w;
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (pool) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
            v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], p.TW));
          }
        }
        if (ok && c0 < p.Cout) {
          if (out_f32) {
            float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) + pix * p.Cout + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
            __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(p.out) + pix * p.Cout + c0;
            for (int pl = 0; pl < P; ++pl) {
              uint32_t w[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
                w[i] = pack_bf16x2(h0, h1);
                v[2 * i] = __fsub_rn(v[2 * i], __bfloat162float(h0));       // exact residual for the next plane
                v[2 * i + 1] = __fsub_rn(v[2 * i + 1], __bfloat162float(h1));
              }
              uint4 *dst = reinterpret_cast<uint4 *>(o + (long long)pl * p.out_plane_stride);
#pragma unroll
              for (int q = 0; q < 4; ++q) dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * a);
      a ^= 1;
      if (a == 0) aph ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * acc_cols);
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode(EncodeTiledFn *out) {
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

static int encode(EncodeTiledFn fn, CUtensorMap *m, void *addr, int rank, const cuuint64_t *dims,
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

// Pick the TH x TW (=128) pixel patch with the fewest wasted lanes.  Pooling needs both rows and
// both columns of every 2x2 window inside one warp: TW in {4, 8, 16}, TH and TW even.
static void pick_patch(int H, int W, bool pool, int *TH, int *TW) {
  const int cands[6][2] = {{8, 16}, {16, 8}, {4, 32}, {2, 64}, {1, 128}, {32, 4}};
  long long best = -1;
  for (auto &c : cands) {
    if (pool && !(c[1] == 16 || c[1] == 8 || c[1] == 4)) continue;
    long long tiles = (long long)ceil_div(H, c[0]) * ceil_div(W, c[1]);
    if (best < 0 || tiles < best) { best = tiles; *TH = c[0]; *TW = c[1]; }
  }
}

static int g_num_sms = 0;

template <int BN>
static int launch_bn(const CUtensorMap &ta, const CUtensorMap &tb, ConvTcParams &p, cudaStream_t st) {
  const size_t stage_bytes = (size_t)p.planes * (kABytes + BN * 128);
  const size_t ctrl = 8 * (2 * kMaxStages + 4) + 16;
  const size_t budget = 227 * 1024;
  int stages = (int)((budget - 1024 - ctrl) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char *e = getenv("CTPN_TC_STAGES")) { int v = atoi(e); if (v >= 1 && v < stages) stages = v; }
  CTPN_REQUIRE(stages >= 1, "conv_tc: a pipeline stage (%zu B) does not fit in shared memory", stage_bytes);
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + ctrl;
  CTPN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  char label[128];
  snprintf(label, sizeof(label), "conv_tc t%d %dx%dx%d c%d-%d p%d bn%d", p.taps, p.B, p.H, p.W, p.Cin, p.Cout, p.planes, BN);
  ProfScope prof(label, 2.0 * p.B * p.H * p.W * (double)p.taps * p.Cin * p.Cout, st);
  conv_tc_kernel<BN><<<grid, kTcThreads, smem, st>>>(ta, tb, p);
  CTPN_LAUNCH_CHECK();
  return CTPN_OK;
}

}  // namespace ctpn

using namespace ctpn;

extern "C" int ctpn_conv3x3(const void *in_planes, const void *w_planes, const float *bias, void *out, int B, int H,
                            int W, int cin, int cout, int taps, int planes, int flags, void *stream) {
  CTPN_REQUIRE(in_planes && w_planes && bias && out, "ctpn_conv3x3: null pointer");
  CTPN_REQUIRE(taps == 9 || taps == 1, "ctpn_conv3x3: taps must be 9 or 1 (got %d)", taps);
  CTPN_REQUIRE(planes >= 1 && planes <= 3, "ctpn_conv3x3: planes must be 1..3 (got %d)", planes);
  CTPN_REQUIRE(cin % 64 == 0 && cin >= 64, "ctpn_conv3x3: Cin must be a multiple of 64 (got %d)", cin);
  CTPN_REQUIRE(cout % 64 == 0 && cout >= 64, "ctpn_conv3x3: Cout must be a multiple of 64 (got %d)", cout);
  CTPN_REQUIRE(B > 0 && H > 0 && W > 0, "ctpn_conv3x3: bad shape");
  const bool pool = flags & CTPN_F_POOL;
  CTPN_REQUIRE(!pool || (taps == 9 && H >= 2 && W >= 2), "ctpn_conv3x3: pooling needs taps=9 and H,W >= 2");
  if (g_num_sms == 0) {
    int dev = 0;
    CTPN_CUDA(cudaGetDevice(&dev));
    CTPN_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  EncodeTiledFn enc = nullptr;
  int rc = get_encode(&enc);
  if (rc) return rc;

  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W; p.Cin = cin; p.Cout = cout; p.taps = taps; p.planes = planes; p.flags = flags;
  pick_patch(H, W, pool, &p.TH, &p.TW);
  p.tw_log2 = 0;
  while ((1 << p.tw_log2) < p.TW) ++p.tw_log2;
  p.tiles_x = ceil_div(W, p.TW);
  p.tiles_y = ceil_div(H, p.TH);
  p.Ho = pool ? H / 2 : H;
  p.Wo = pool ? W / 2 : W;
  p.bias = bias;
  p.out = out;
  p.out_plane_stride = (long long)B * p.Ho * p.Wo * cout;
  p.cout_pad = cout;

  int BN = planes == 1 ? 256 : 128;
  if (const char *e = getenv("CTPN_TC_BN")) { int v = atoi(e); if (v == 64 || v == 128 || v == 256) BN = v; }
  if (planes > 1 && BN > 128) BN = 128;   // two accumulators per tile buffer: 4 * BN TMEM columns <= 512
  while (BN > cout || cout % BN) BN >>= 1;
  p.tiles_n = cout / BN;
  const long long total = (long long)B * p.tiles_x * p.tiles_y * p.tiles_n;
  CTPN_REQUIRE(total < (1ll << 31), "ctpn_conv3x3: too many tiles");
  p.total_tiles = (int)total;

  CUtensorMap ta, tb;
  {
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes * B};
    cuuint64_t strides[3] = {(cuuint64_t)cin * 2, (cuuint64_t)W * cin * 2, (cuuint64_t)H * W * cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
    if ((rc = encode(enc, &ta, const_cast<void *>(in_planes), 4, dims, strides, box))) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)taps * cin, (cuuint64_t)planes * cout};
    cuuint64_t strides[1] = {(cuuint64_t)taps * cin * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    if ((rc = encode(enc, &tb, const_cast<void *>(w_planes), 2, dims, strides, box))) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (BN) {
    case 256: return launch_bn<256>(ta, tb, p, st);
    case 128: return launch_bn<128>(ta, tb, p, st);
    default: return launch_bn<64>(ta, tb, p, st);
  }
}
