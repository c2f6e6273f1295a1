// CTA-pair tcgen05 GEMM (v3): a cluster of two CTAs (two SMs of one TPC) computes a 256 x BN output tile with
// tcgen05.mma.cta_group::2 (UMMA M = 256, N = BN, K = 16).
//
// Why: with one SM per tile the UMMA reads A and B from that SM's shared memory while TMA writes the next stage into it; at
// 128 B/clk the 128x128 tile is shared-memory bound at ~50% of the tensor peak and 128x256 at ~68% (measured: DESIGN.md §3).  In
// pair mode each SM stages its own 128 A rows and only HALF of the B tile (BN/2 rows); the tensor cores of both SMs consume
// both halves, so per SM the smem traffic per FLOP halves.
//
//   rank r of the pair : TMA-loads A rows [m0 + 128 r, +128) and B rows [n0 + r BN/2, + BN/2) into ITS smem
//                        (cp.async.bulk.tensor ... .cta_group::2, completion on the LEADER's full barrier)
//   leader (rank 0)    : one thread issues tcgen05.mma.cta_group::2; tcgen05.commit multicasts the "stage free" /
//                        "accumulator ready" arrivals to the barriers of both CTAs
//   both CTAs          : epilogue of their own 128 accumulator rows (TMEM lanes) exactly like the single-CTA kernel
#include "gemm_tc.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace dvd {
using namespace tc;

constexpr int QBM = 128, QBK = 64;       // rows per CTA (256 per pair), K per stage

template <int BN>
struct QCfg {
  static constexpr int A_BYTES = QBM * QBK * 2, B_BYTES = (BN / 2) * QBK * 2;       // per CTA
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : 4;       // <= 96 KB ring: two CTAs (of different pairs) per SM overlap epilogue and main loop
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int CW = 128;
  static constexpr int SLD = CW + 4;
  static constexpr int STAGING_BYTES = 4 * 32 * SLD * 4;
  static_assert(STAGING_BYTES <= RING_BYTES, "epilogue staging reuses the drained ring");
  static constexpr int SMEM = RING_BYTES + 1024 + 256;
  static_assert(SMEM <= 232448, "shared memory budget");
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load for a CTA pair: data lands in THIS CTA's smem, the transaction bytes are counted on the barrier at `bar_cluster_addr`
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs have completed) on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 2)
k_gemm_tc3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, Epilogue e) {
  using Cfg = QCfg<BN>;
  constexpr int ST = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES);
  uint64_t* empty = full + ST;
  uint64_t* tmem_full = empty + ST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader
  const int tile_n = blockIdx.x >> 1;
  const int m0 = blockIdx.y * 256 + (int)rank * QBM;       // this CTA's 128 rows
  const int n0 = tile_n * BN;
  const int nkb = (K + QBK - 1) / QBK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }     // full: one arrival per CTA of the pair
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, BN);
  fence_before_sync();
  cluster_sync_all();                                      // barriers of BOTH CTAs are initialised before any remote arrive / TMA
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own A rows, own half of the B tile; bytes are counted on the leader's full barrier
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % ST, it = kb / ST;
        mbar_wait(&empty[s], (it & 1) ^ 1);
        uint8_t* a = smem + s * Cfg::STAGE_BYTES;
        const uint32_t lead_full = mapa(smem_u32(&full[s]), 0);
        if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);        // both CTAs' loads of this stage
        else mbar_arrive_remote(lead_full);
        tma_load_2d_pair(a, &tmA, lead_full, kb * QBK, m0);
        tma_load_2d_pair(a + Cfg::A_BYTES, &tmB, lead_full, kb * QBK, n0 + (int)rank * (BN / 2));
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank == 0) {
    if (lane == 0) {
      // ===== MMA issuer (leader only): UMMA M = 256 across the pair
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % ST, it = kb / ST;
        mbar_wait(&full[s], it & 1);
        fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES), b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < QBK / 16; ++k)
          mma_f16_ss_pair(tmem_base, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc, (kb | k) ? 1u : 0u);
        mma_commit_pair(&empty[s]);                          // frees stage s in both CTAs
      }
      mma_commit_pair(tmem_full);                            // accumulators of both CTAs are complete
    }
    __syncwarp();
  }

  // ===== epilogue (both CTAs, own 128 rows); identical to k_gemm_tc
  mbar_wait(tmem_full, 0);
  fence_after_sync();
  constexpr int CW = Cfg::CW, SLD = Cfg::SLD;
  float* stage = reinterpret_cast<float*>(smem) + warp * 32 * SLD;
  const int row_t = m0 + warp * 32 + lane;
#pragma unroll 1
  for (int pass = 0; pass < BN / CW; ++pass) {
#pragma unroll 1
    for (int c0 = 0; c0 < CW; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(pass * CW + c0), r);
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
    __syncwarp();
    const int col = n0 + pass * CW + 4 * lane;
    if (col < N) {
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
            const int row = m0 + warp * 32 + r0 + i;
            const int rr = e.resid_mod ? (row % e.resid_mod) : row;
            q[i] = *reinterpret_cast<const float4*>(e.resid + (size_t)rr * e.ldr + col);
          }
        }
        if (e.pos) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = m0 + warp * 32 + r0 + i;
            p[i] = __ldg(reinterpret_cast<const float4*>(e.pos + (size_t)(row % e.pos_rows) * N + col));
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = m0 + warp * 32 + r0 + i;
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
    __syncwarp();
  }
  fence_before_sync();
  cluster_sync_all();                                        // the peer's smem / TMEM must outlive every MMA that reads it
  if (warp == 1) tmem_dealloc2(tmem_base, BN);
}

template <int BN>
static int launch3(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const Epilogue& e, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DVD_CUDA(cudaFuncSetAttribute(k_gemm_tc3<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, QCfg<BN>::SMEM));
    attr_set = true;
  }
  dim3 grid(2 * cdiv(N, BN), M / 256);
  k_gemm_tc3<BN><<<grid, 128, QCfg<BN>::SMEM, st>>>(tmA, tmB, M, N, K, e);
  DVD_LAUNCH_CHECK("k_gemm_tc3");
  return 0;
}

// M % 256 == 0, N % BN == 0 (BN = 256 or 128) required; the caller falls back to the single-CTA kernels otherwise.
int gemm_tc3_dispatch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K, const Epilogue& e, int bn,
                      cudaStream_t st) {
  DVD_REQUIRE(M % 256 == 0 && (bn == 128 || bn == 256) && N % bn == 0, "gemm_tc3: unsupported shape M=%d N=%d bn=%d", M, N, bn);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64); if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(bn / 2), 64); if (rc) return rc;
  return bn == 256 ? launch3<256>(tmA, tmB, M, N, K, e, st) : launch3<128>(tmA, tmB, M, N, K, e, st);
}

}  // namespace dvd
