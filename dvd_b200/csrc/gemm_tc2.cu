// Persistent tcgen05 GEMM (v2): one CTA per SM loops over 128 x BN output tiles; the accumulator is double-buffered in
// tensor memory so that the epilogue of tile i overlaps the main loop of tile i+1.
//
//   warp 0 (one lane)  TMA producer : ring of STAGES x (A 128x64 + B BNx64 bf16, SWIZZLE_128B), continuous across tiles
//   warp 1 (one lane)  MMA issuer   : UMMA 128 x BN x 16, accumulator buffer = tile parity; tcgen05.commit -> empty[s] / acc_full[buf]
//   warps 2..5         epilogue     : tcgen05.ld (TMEM lane quarter = warp % 4) -> private fp32 staging tile in smem ->
//                                     fused Epilogue with fully coalesced row segments; acc_empty[buf] is released as soon as
//                                     the accumulator has been copied out of TMEM
//   BN in {64, 128, 192, 256} is chosen per problem so that the number of tiles is close to a multiple of the SM count
//   (M = 2048 token GEMMs have only 96..576 tiles): see pick_bn().
#include "gemm_tc.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace dvd {
using namespace tc;

constexpr int PBM = 128, PBK = 64;

template <int BN>
struct PCfg {
  static constexpr int A_BYTES = PBM * PBK * 2, B_BYTES = BN * PBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : ((BN == 64) ? 6 : 4);
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int CW = (BN % 128 == 0) ? 128 : 64;                   // columns per epilogue pass
  static constexpr int SLD = CW + 4;                                      // fp32 staging row stride
  static constexpr int STAGING_BYTES = 4 * 32 * SLD * 4;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512);
  static constexpr int SMEM = RING_BYTES + STAGING_BYTES + 1024 + 256;
  static_assert(SMEM <= 232448, "shared memory budget");
};

struct ConvGeom2 { int H, W, Cin; };

template <int BN, bool CONV>
__global__ void __launch_bounds__(192, 1) k_gemm_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                     int M, int N, int K, int tiles_n, int num_tiles, Epilogue e, ConvGeom2 cg) {
  using Cfg = PCfg<BN>;
  constexpr int ST = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* staging = reinterpret_cast<float*>(smem + Cfg::RING_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES + Cfg::STAGING_BYTES);
  uint64_t* empty = full + ST;
  uint64_t* acc_full = empty + ST;      // 2
  uint64_t* acc_empty = acc_full + 2;   // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + PBK - 1) / PBK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      uint32_t g = 0;                                               // global k-block counter (ring position)
      const int cblocks = CONV ? cg.Cin / 64 : 1;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * PBM, n0 = (tile % tiles_n) * BN;
        int cn = 0, cy = 0, cx = 0;
        if (CONV) {
          const int hw = cg.H * cg.W;
          cn = m0 / hw; const int rem = m0 % hw; cy = rem / cg.W; cx = rem % cg.W;   // 128 consecutive pixels of one image row
        }
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % ST;
          mbar_wait(&empty[s], ((g / ST) & 1) ^ 1);
          uint8_t* a = smem + s * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          if (CONV) {
            const int tap = kb / cblocks, cb = kb % cblocks;
            tma_load_4d(a, &tmA, &full[s], cb * 64, cx + tap % 3 - 1, cy + tap / 3 - 1, cn);
          } else {
            tma_load_2d(a, &tmA, &full[s], kb * PBK, m0);
          }
          tma_load_2d(a + Cfg::A_BYTES, &tmB, &full[s], kb * PBK, n0);                 // one box of BN rows (BN <= 256)
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(PBM, BN);
      uint32_t g = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);            // epilogue has drained this accumulator buffer
        fence_after_sync();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % ST;
          mbar_wait(&full[s], (g / ST) & 1);
          fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES), b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < PBK / 16; ++k)
            mma_f16_ss(tacc, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc, (kb | k) ? 1u : 0u);
          mma_commit(&empty[s]);
        }
        mma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps (2..5): TMEM lane quarter = warp % 4
    const int quarter = warp & 3;
    constexpr int CW = Cfg::CW, SLD = Cfg::SLD;
    float* stage = staging + quarter * 32 * SLD;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int m0 = (tile / tiles_n) * PBM, n0 = (tile % tiles_n) * BN;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      fence_after_sync();
      const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(quarter * 32) << 16);
      const int row_t = m0 + quarter * 32 + lane;
#pragma unroll 1
      for (int pass = 0; pass < BN / CW; ++pass) {
        // ---- phase 1: TMEM -> registers -> staging (+ transposed V^T store, coalesced in this mapping)
#pragma unroll 1
        for (int c0 = 0; c0 < CW; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tacc + (uint32_t)(pass * CW + c0), r);
          tmem_ld_wait();
          float* srow = stage + lane * SLD + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(srow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                               __uint_as_float(r[j + 3]));
          const int col0 = n0 + pass * CW + c0;
          if (e.vt_out && col0 >= e.vt_col0 && col0 + 31 < N) {
            float bv[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + col0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
              bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
            }
            __nv_bfloat16* o = e.vt_out + ((size_t)(row_t >> 10) * (N - e.vt_col0) + (col0 - e.vt_col0)) * 1024 + (row_t & 1023);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[(size_t)j * 1024] = __float2bfloat16_rn(__uint_as_float(r[j]) + bv[j]);
          }
        }
        if (pass == BN / CW - 1) {                                   // accumulator fully copied out: hand the buffer back to the MMA warp
          fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
        }
        __syncwarp();
        // ---- phase 2: staging -> fused epilogue -> coalesced global stores (lane = 4 consecutive columns)
        const int col = n0 + pass * CW + 4 * lane;
        if (4 * lane < CW && col < N) {
          float4 cb = make_float4(0.f, 0.f, 0.f, 0.f), cs = make_float4(1.f, 1.f, 1.f, 1.f), ct = cb, cgate = cs;
          if (e.bias) cb = __ldg(reinterpret_cast<const float4*>(e.bias + col));
          if (e.scale) { cs = __ldg(reinterpret_cast<const float4*>(e.scale + col)); ct = __ldg(reinterpret_cast<const float4*>(e.shift + col)); }
          if (e.gate) cgate = __ldg(reinterpret_cast<const float4*>(e.gate + col));
          const bool has_scale = e.scale != nullptr, has_gate = e.gate != nullptr;
          const int act = e.act;
#pragma unroll 1
          for (int r0 = 0; r0 < 32; r0 += 8) {
            float4 a[8], q[8], p[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(stage + (r0 + i) * SLD + 4 * lane);
            if (e.resid) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = m0 + quarter * 32 + r0 + i;
                const int rr = e.resid_mod ? (row % e.resid_mod) : row;
                q[i] = *reinterpret_cast<const float4*>(e.resid + (size_t)rr * e.ldr + col);      // may alias e.out (in-place residual)
              }
            }
            if (e.pos) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = m0 + quarter * 32 + r0 + i;
                p[i] = __ldg(reinterpret_cast<const float4*>(e.pos + (size_t)(row % e.pos_rows) * N + col));
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = m0 + quarter * 32 + r0 + i;
              float v[4] = {a[i].x + cb.x, a[i].y + cb.y, a[i].z + cb.z, a[i].w + cb.w};
              if (has_scale) { v[0] = v[0] * cs.x + ct.x; v[1] = v[1] * cs.y + ct.y; v[2] = v[2] * cs.z + ct.z; v[3] = v[3] * cs.w + ct.w; }
              if (act == ACT_RELU) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
              else if (act == ACT_GELU) { v[0] = gelu_tanh_fast(v[0]); v[1] = gelu_tanh_fast(v[1]); v[2] = gelu_tanh_fast(v[2]); v[3] = gelu_tanh_fast(v[3]); }
              else if (act == ACT_SIGMOID) { v[0] = sigmoidf_(v[0]); v[1] = sigmoidf_(v[1]); v[2] = sigmoidf_(v[2]); v[3] = sigmoidf_(v[3]); }
              if (e.pos) { v[0] += p[i].x; v[1] += p[i].y; v[2] += p[i].z; v[3] += p[i].w; }
              if (has_gate) { v[0] *= cgate.x; v[1] *= cgate.y; v[2] *= cgate.z; v[3] *= cgate.w; }
              if (e.resid) { v[0] += q[i].x; v[1] += q[i].y; v[2] += q[i].z; v[3] += q[i].w; }
              int orow, ocol;
              epilogue_dest(e, row, col, orow, ocol);
              if (e.out) *reinterpret_cast<float4*>(e.out + (size_t)orow * e.ldc + ocol) = make_float4(v[0], v[1], v[2], v[3]);
              if (e.out_bf16) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
                *reinterpret_cast<uint2*>(e.out_bf16 + (size_t)orow * e.ldc_bf16 + ocol) = u;
              }
            }
          }
        }
        __syncwarp();                                                // staging tile is reused by the next pass / tile
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// cycles per k-block of one CTA (UMMA time vs shared-memory traffic of TMA writes + UMMA reads at 128 B/clk), see DESIGN.md §3
static int kblock_cost(int bn) { return bn == 64 ? 384 : (bn == 128 ? 512 : (bn == 192 ? 640 : 768)); }

int pick_bn(int M, int N) {
  const int mt = (M + PBM - 1) / PBM;
  int best = 128; long long best_cost = -1;
  for (int bn : {64, 128, 192, 256}) {
    if (bn > 64 && N < bn / 2) continue;
    const long long tiles = (long long)mt * ((N + bn - 1) / bn);
    const long long waves = (tiles + kSMs - 1) / kSMs;
    const long long cost = waves * kblock_cost(bn);
    if (best_cost < 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

template <int BN, bool CONV>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const Epilogue& e, ConvGeom2 cg, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DVD_CUDA(cudaFuncSetAttribute(k_gemm_tc2<BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<BN>::SMEM));
    attr_set = true;
  }
  const int tiles_n = cdiv(N, BN), num_tiles = cdiv(M, PBM) * tiles_n;
  const int grid = num_tiles < kSMs ? num_tiles : kSMs;
  k_gemm_tc2<BN, CONV><<<grid, 192, PCfg<BN>::SMEM, st>>>(tmA, tmB, M, N, K, tiles_n, num_tiles, e, cg);
  DVD_LAUNCH_CHECK("k_gemm_tc2");
  return 0;
}

// A: 2-D [M,K] (lda) or, when conv_h > 0, NHWC [B,H,W,Cin] with K = 9*Cin; W: [N,K] K-major.
int gemm_tc2_dispatch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K, const Epilogue& e,
                      int conv_b, int conv_h, int conv_w, int conv_cin, cudaStream_t st) {
  const bool conv = conv_h > 0;
  int bn = pick_bn(M, N);
  if (const char* f = getenv("DVD_GEMM_BN")) { int v = atoi(f); if (v == 64 || v == 128 || v == 192 || v == 256) bn = v; }   // tuning override
  CUtensorMap tmA, tmB;
  int rc;
  if (conv) rc = make_tmap_bf16_nhwc(&tmA, A, (uint64_t)conv_b, (uint64_t)conv_h, (uint64_t)conv_w, (uint64_t)conv_cin);
  else rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)bn, 64);
  if (rc) return rc;
  ConvGeom2 cg{conv_h, conv_w, conv_cin};
#define DVD_TC2(BNv)                                                                     \
  if (bn == BNv) return conv ? launch2<BNv, true>(tmA, tmB, M, N, K, e, cg, st) : launch2<BNv, false>(tmA, tmB, M, N, K, e, cg, st);
  DVD_TC2(64) DVD_TC2(128) DVD_TC2(192) DVD_TC2(256)
#undef DVD_TC2
  set_error("gemm_tc2: no tile configuration"); return DVD_E_BADARG;
}

}  // namespace dvd
