// Persistent CTA-pair tcgen05 GEMM: the kernel every dense layer of the denoiser runs on.
//
//   C[M,N] = epilogue(A[M,K] * W[N,K]^T)      A, W K-major 16-bit (bf16, or bf16 hi + lo pairs = 3 tensor-core passes, DVD_PREC_BF16X3)
//
// Shape of the machine: 74 clusters of two CTAs (the two SMs of a TPC), one cluster per TPC, each looping over work units.
//   * a unit = one 256 x BN output tile (UMMA M = 256 across the pair, tcgen05.mma.cta_group::2) x one K range (split-K);
//     units are dealt round-robin (unit u -> pair u % npairs), so that the M = 2048 problems of a single document, which have
//     only 48..144 tiles, still load every SM: the host picks BN in {64,128,192,256} and the split count from a cycle model.
//   * per SM the pair mode stages 128 A rows and only HALF of the W tile (BN/2 rows): 32 KB instead of 48 KB of L2 -> SM traffic per
//     k-block of a 128 x 256 tile, which is what bounds a one-SM-per-tile kernel (~13 TB/s of L2 reads at 1.1 PFLOP/s).
//   * warp 0 (one lane, both CTAs)  TMA producer: ring of 3..8 stages of [A_hi | A_lo | W_hi | W_lo] boxes (SWIZZLE_128B), continuous
//     across units; transaction bytes of both CTAs are counted on the LEADER's full barrier.
//   * warp 1 (one lane, leader)     MMA issuer: 4 (or 12: hi*hi, lo*hi, hi*lo) UMMA 256 x BN x 16 per stage into one of TWO TMEM
//     accumulators; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs.
//   * warps 2..9 (both CTAs)        epilogue of the CTA's own 128 rows, overlapping the next unit's main loop: each warp owns a TMEM
//     lane quarter x half of the columns; tcgen05.ld 32x32 -> private XOR-swizzled 4 KB staging tile -> row-coalesced fused epilogue
//     (8 lanes per 128-byte row segment).
//   * split-K: the units of a tile with split < splits-1 store their raw accumulators to an fp32 slice of the scratch and bump the
//     tile's arrival counter; the LAST split (highest unit index) waits for the counter, adds the slices in index order (bitwise
//     deterministic) and runs the fused epilogue.  A pair handles its units in increasing order and a unit only ever waits for units
//     with a lower index, all clusters are co-resident (grid <= cudaOccupancyMaxActiveClusters), so the wait cannot deadlock.
//   * CONV: A is an NHWC activation read through a 4-D tensor map (implicit GEMM of the 3x3 pyramid convolutions, zero padding = TMA
//     out-of-bounds fill); a pair covers 256 consecutive pixels of one image row.
#include "gemm_tc.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"
#include <stdlib.h>

namespace dvd {
using namespace tc;

constexpr int PBM = 128, PBK = 64;
constexpr int PP_THREADS = 320;           // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int PP_EPI_WARPS = 8;

template <int BN, bool X3>
struct PairCfg {
  static constexpr int NOP = X3 ? 2 : 1;
  static constexpr int A_BYTES = PBM * PBK * 2;                 // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * PBK * 2;            // this CTA's half of the W tile
  static constexpr int STAGE_BYTES = NOP * (A_BYTES + B_BYTES);
  static constexpr int B_OFF = NOP * A_BYTES;
  static constexpr int STAGING_BYTES = PP_EPI_WARPS * 32 * 32 * 4;   // one 32 x 32 fp32 tile per epilogue warp
  static constexpr int FIXED = STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int STAGES_FIT = (232448 - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512);
  static constexpr int SMEM = RING_BYTES + FIXED;
  static_assert(STAGES >= 2 && SMEM <= 232448, "shared memory budget");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "SWIZZLE_128B tiles are 1024-byte aligned");
};

struct PairParams {
  int M, N, K;
  int tiles_n, splits, units, npairs;
  int conv_h, conv_w, conv_cin;
  float* partial;                 // [(splits-1)][M][N] fp32
  unsigned int* counters;         // [tiles]
};

// ---- cluster / pair PTX helpers
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the data lands in THIS CTA's smem, the transaction bytes are counted on the barrier at `bar_cluster_addr`
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct PairUnit { int tile, split, m0, n0, kb0, kb1; };
__device__ __forceinline__ PairUnit decode_unit(const PairParams& p, int u, int nkb, int bn) {
  PairUnit q;
  q.tile = u / p.splits; q.split = u - q.tile * p.splits;
  const int mt = q.tile / p.tiles_n, nt = q.tile - mt * p.tiles_n;
  q.m0 = mt * (2 * PBM); q.n0 = nt * bn;
  q.kb0 = (int)(((long long)q.split * nkb) / p.splits);
  q.kb1 = (int)(((long long)(q.split + 1) * nkb) / p.splits);
  return q;
}

template <int BN, bool X3, bool CONV>
__global__ void __launch_bounds__(PP_THREADS, 1)
k_gemm_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmBl, const PairParams p, const Epilogue e) {
  using Cfg = PairCfg<BN, X3>;
  constexpr int ST = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* staging = reinterpret_cast<float*>(smem + Cfg::RING_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES + Cfg::STAGING_BYTES);
  uint64_t* empty = full + ST;
  uint64_t* acc_full = empty + ST;        // 2
  uint64_t* acc_empty = acc_full + 2;     // 2 (the leader's are used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                   // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int nkb = (p.K + PBK - 1) / PBK;

  pdl_trigger();
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    if (X3) { prefetch_tmap(&tmAl); prefetch_tmap(&tmBl); }
    for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }       // full: one arrival per CTA of the pair
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 2 * PP_EPI_WARPS); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
  fence_before_sync();
  cluster_sync_all();                                        // the barriers of BOTH CTAs exist before any remote arrive / TMA
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                                // the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own 128 A rows, own half of the W tile; bytes are counted on the leader's full barrier
      uint32_t g = 0;
      const int cblocks = CONV ? p.conv_cin / 64 : 1;
      for (int u = pair; u < p.units; u += p.npairs) {
        const PairUnit q = decode_unit(p, u, nkb, BN);
        const int m0 = q.m0 + (int)rank * PBM, nb = q.n0 + (int)rank * (BN / 2);
        int cn = 0, cy = 0, cx = 0;
        if (CONV) {
          const int hw = p.conv_h * p.conv_w;
          cn = m0 / hw; const int rem = m0 - cn * hw; cy = rem / p.conv_w; cx = rem - cy * p.conv_w;   // 128 consecutive pixels of one image row
        }
        for (int kb = q.kb0; kb < q.kb1; ++kb, ++g) {
          const int s = g % ST;
          mbar_wait(&empty[s], ((g / ST) & 1) ^ 1);
          uint8_t* a = smem + s * Cfg::STAGE_BYTES;
          const uint32_t lead_full = mapa(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);        // both CTAs' loads of this stage
          else mbar_arrive_cluster(lead_full);
#pragma unroll
          for (int o = 0; o < Cfg::NOP; ++o) {
            const CUtensorMap* ta = o ? &tmAl : &tmA;
            const CUtensorMap* tb = o ? &tmBl : &tmB;
            if (CONV) {
              const int tap = kb / cblocks, cb = kb - tap * cblocks;
              tma_load_4d_pair(a + o * Cfg::A_BYTES, ta, lead_full, cb * 64, cx + tap % 3 - 1, cy + tap / 3 - 1, cn);
            } else {
              tma_load_2d_pair(a + o * Cfg::A_BYTES, ta, lead_full, kb * PBK, m0);
            }
            tma_load_2d_pair(a + Cfg::B_OFF + o * Cfg::B_BYTES, tb, lead_full, kb * PBK, nb);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      // ===== MMA issuer (leader only): UMMA M = 256 across the pair, accumulator buffer = unit parity
      constexpr uint32_t idesc = make_idesc_bf16(2 * PBM, BN);
      uint32_t g = 0;
      int it = 0;
      for (int u = pair; u < p.units; u += p.npairs, ++it) {
        const PairUnit q = decode_unit(p, u, nkb, BN);
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);       // both CTAs' epilogue warps have drained this accumulator
        fence_after_sync();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = q.kb0; kb < q.kb1; ++kb, ++g) {
          const int s = g % ST;
          mbar_wait(&full[s], (g / ST) & 1);
          fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES), b_addr = a_addr + Cfg::B_OFF;
#pragma unroll
          for (int k = 0; k < PBK / 16; ++k) {
            const uint64_t ah = make_desc_k_sw128(a_addr + k * 32), bh = make_desc_k_sw128(b_addr + k * 32);
            mma_f16_ss_pair(tacc, ah, bh, idesc, (kb > q.kb0 || k > 0) ? 1u : 0u);
            if (X3) {
              mma_f16_ss_pair(tacc, make_desc_k_sw128(a_addr + Cfg::A_BYTES + k * 32), bh, idesc, 1u);     // lo * hi
              mma_f16_ss_pair(tacc, ah, make_desc_k_sw128(b_addr + Cfg::B_BYTES + k * 32), idesc, 1u);     // hi * lo
            }
          }
          mma_commit_pair(&empty[s]);                          // frees stage s in both CTAs
        }
        mma_commit_pair(&acc_full[buf]);                       // the accumulators of both CTAs are complete
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps (both CTAs): TMEM lane quarter = warp % 4, column half = (warp - 2) / 4
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    float* stage = staging + (warp - 2) * (32 * 32);
    const uint32_t lead_acc_empty0 = mapa(smem_u32(&acc_empty[0]), 0);
    const int prow = lane >> 3, pc = lane & 7;                 // phase 2: 4 rows x 8 column quads per warp instruction
    constexpr int NCH = BN / 64;                               // 32-column chunks per warp and unit
    int it = 0;
    for (int u = pair; u < p.units; u += p.npairs, ++it) {
      const PairUnit q = decode_unit(p, u, nkb, BN);
      const int buf = it & 1;
      const int rbase = q.m0 + (int)rank * PBM + quarter * 32;  // first global row of this warp
      const bool reducer = q.split == p.splits - 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      fence_after_sync();
      const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * (BN / 2));
      bool waited = false;
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        const int col0 = q.n0 + half * (BN / 2) + ch * 32;
        // ---- phase 1 (thread = row = TMEM lane): TMEM -> registers -> swizzled staging (+ transposed V^T store, coalesced in this mapping)
        uint32_t r[32];
        tmem_ld_32x32(tacc + (uint32_t)(ch * 32), r);
        tmem_ld_wait();
        if (ch == NCH - 1) {                                    // accumulator fully copied out: hand the buffer back to the MMA warp
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead_acc_empty0 + (uint32_t)buf * 8u);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        if (e.vt_out && col0 >= e.vt_col0) {                    // splits == 1 (host), bias-only epilogue
          float bv[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + col0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
          }
          const int row_t = rbase + lane;
          uint16_t* o = reinterpret_cast<uint16_t*>(e.vt_out) + ((size_t)(row_t >> 10) * (p.N - e.vt_col0) + (col0 - e.vt_col0)) * 1024 + (row_t & 1023);
#pragma unroll
          for (int j = 0; j < 32; ++j) o[(size_t)j * 1024] = cvt16(__uint_as_float(r[j]) + bv[j], e.out_f16);
        }
        __syncwarp();
        // ---- phase 2 (8 lanes = one 128-byte row segment): staging -> [split-K exchange] -> fused epilogue -> global
        const int col = col0 + 4 * pc;
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + prow;
          a[i] = *reinterpret_cast<const float4*>(stage + rr * 32 + ((pc ^ (rr & 7)) << 2));
        }
        if (p.splits > 1 && !reducer) {
          float* ps = p.partial + ((size_t)q.split * p.M + rbase) * p.N + col;
#pragma unroll
          for (int i = 0; i < 8; ++i) __stcg(reinterpret_cast<float4*>(ps + (size_t)(i * 4 + prow) * p.N), a[i]);
        } else {
          if (p.splits > 1) {
            if (!waited) {                                      // every lower split of this tile has stored its slice
              if (lane == 0) {
                const unsigned int target = (unsigned int)(2 * PP_EPI_WARPS * (p.splits - 1));
                for (uint32_t spins = 0; ld_acquire_u32(p.counters + q.tile) < target; ++spins) {
                  __nanosleep(40);
                  if (spins > (1u << 24)) __trap();
                }
              }
              __syncwarp();
              waited = true;
            }
            for (int s2 = 0; s2 < p.splits - 1; ++s2) {
              const float* ps = p.partial + ((size_t)s2 * p.M + rbase) * p.N + col;
              float4 t[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) t[i] = __ldcg(reinterpret_cast<const float4*>(ps + (size_t)(i * 4 + prow) * p.N));
#pragma unroll
              for (int i = 0; i < 8; ++i) { a[i].x += t[i].x; a[i].y += t[i].y; a[i].z += t[i].z; a[i].w += t[i].w; }
            }
          }
          const EpiCols ec = load_epi_cols(e, col);
          float4 qv[8], pv[8];
          if (e.resid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = rbase + i * 4 + prow;
              const int rr = e.resid_mod ? (row % e.resid_mod) : row;
              qv[i] = *reinterpret_cast<const float4*>(e.resid + (size_t)rr * e.ldr + col);      // may alias e.out (in-place residual)
            }
          }
          if (e.pos) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = rbase + i * 4 + prow;
              pv[i] = __ldg(reinterpret_cast<const float4*>(e.pos + (size_t)(row % e.pos_rows) * p.N + col));
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = rbase + i * 4 + prow;
            float v[4];
            apply_epi4(ec, a[i], e.pos != nullptr, pv[i], e.resid != nullptr, qv[i], v);
            int orow, ocol;
            epilogue_dest(e, row, col, orow, ocol);
            store_tc_out4(e, orow, ocol, v);
          }
        }
        __syncwarp();                                           // the staging tile is rewritten by the next chunk
      }
      if (p.splits > 1) {
        if (!reducer) __threadfence();                          // slice stores visible before the arrival
        __syncwarp();
        if (lane == 0) {
          const unsigned int old = atomicAdd(p.counters + q.tile, 1u);
          // the last of the 16 reducer warps leaves the counter at zero for the next launch
          if (reducer && old == (unsigned int)(2 * PP_EPI_WARPS * p.splits - 1)) atomicExch(p.counters + q.tile, 0u);
        }
      }
    }
  }
  fence_before_sync();
  cluster_sync_all();                                          // the peer's smem / TMEM must outlive every MMA that reads it
  if (warp == 1) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------- host side
static int g_force_bn = -1, g_force_splits = -1;
static void read_env() {
  if (g_force_bn >= 0) return;
  const char* b = getenv("DVD_GEMM_BN"); g_force_bn = b ? atoi(b) : 0;
  const char* s = getenv("DVD_GEMM_SPLITS"); g_force_splits = s ? atoi(s) : 0;
}

bool gemm_pair_supported(int M, int N, int K, bool conv) {
  (void)K; (void)conv;
  return M % 256 == 0 && N % 64 == 0;
}

template <int BN, bool X3, bool CONV>
static int max_pairs() {           // co-resident clusters of this instantiation on the current device (cached per device)
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev]) return cached[dev];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 256); cfg.blockDim = dim3(PP_THREADS); cfg.dynamicSmemBytes = PairCfg<BN, X3>::SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_gemm_pair<BN, X3, CONV>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = sm_count() / 2 - 2; }
  if (n > sm_count() / 2) n = sm_count() / 2;
  cached[dev] = n;
  return n;
}

// Cycle model used to pick the tile width and the split count.  Per k-block and CTA: tensor time (UMMA 256 x bn x 16 on a pair =
// bn/2 clocks, 4 per k-block, x3 in split-precision mode) against the L2 -> SM operand stream (~45 B/clk per SM when every SM pulls).
static double unit_kb_cycles(int bn, bool x3) {
  const double mma = (x3 ? 3.0 : 1.0) * 4.0 * (bn / 2.0);
  const double bytes = (x3 ? 2.0 : 1.0) * (16384.0 + bn * 64.0);
  const double mem = bytes / 45.0;
  return mma > mem ? mma : mem;
}

static void pick_config(int M, int N, int K, bool x3, bool can_split, const TcScratch* sk, int npairs, int& bn_out, int& sp_out) {
  const int nkb = (K + PBK - 1) / PBK;
  double best = -1.0;
  bn_out = 64; sp_out = 1;
  const int bns[4] = {256, 192, 128, 64};
  const int sps[6] = {1, 2, 3, 4, 6, 8};
  for (int bi = 0; bi < 4; ++bi) {
    const int bn = bns[bi];
    if (N % bn) continue;
    if (g_force_bn > 0 && bn != g_force_bn && N % g_force_bn == 0) continue;
    const long long tiles = (long long)(M / 256) * (N / bn);
    for (int si = 0; si < 6; ++si) {
      const int sp = sps[si];
      if (sp > 1) {
        if (!can_split || !sk || !sk->partial || !sk->counters) break;
        if (nkb / sp < 4) break;                                             // keep >= 4 k-blocks per unit
        if (tiles > sk->n_counters || (size_t)(sp - 1) * M * N > sk->partial_floats) break;
        if (tiles * sp > 4LL * npairs) break;                                // splitting only pays while the machine is under-filled
      }
      if (g_force_splits > 0 && sp != g_force_splits && sp != 1) continue;
      const long long units = tiles * sp;
      const long long waves = (units + npairs - 1) / npairs;
      const double kbs = (double)((nkb + sp - 1) / sp);
      // per unit: main loop + pipeline fill; per launch: the last epilogue (exposed) + the split exchange
      const double cost = waves * (kbs * unit_kb_cycles(bn, x3) + 700.0) + (900.0 + 9.0 * bn) + (sp > 1 ? 1500.0 + 2.0 * bn * (sp - 1) : 0.0);
      if (best < 0 || cost < best) { best = cost; bn_out = bn; sp_out = sp; }
    }
  }
  if (g_force_splits > 0 && can_split && sk && sk->partial && sp_out != g_force_splits) {
    const long long tiles = (long long)(M / 256) * (N / bn_out);
    if (nkb / g_force_splits >= 1 && tiles <= sk->n_counters && (size_t)(g_force_splits - 1) * M * N <= sk->partial_floats) sp_out = g_force_splits;
  }
}

template <int BN, bool X3, bool CONV>
static int launch_pair(const TcMat& A, const TcMat& W, int M, int N, int K, const Epilogue& e, int conv_b, int conv_h, int conv_w, int conv_cin,
                       int splits, const TcScratch* sk, cudaStream_t st) {
  using Cfg = PairCfg<BN, X3>;
  auto kern = k_gemm_pair<BN, X3, CONV>;
  DVD_SET_MAX_SMEM(kern, Cfg::SMEM);
  const int npairs_max = max_pairs<BN, X3, CONV>();
  DVD_REQUIRE(npairs_max > 0, "gemm_pair: no co-resident cluster available");
  CUtensorMap tmA, tmB, tmAl, tmBl;
  int rc;
  if (CONV) rc = make_tmap_bf16_nhwc(&tmA, A.hi, (uint64_t)conv_b, (uint64_t)conv_h, (uint64_t)conv_w, (uint64_t)conv_cin);
  else rc = make_tmap_bf16_2d(&tmA, A.hi, (uint64_t)M, (uint64_t)K, (uint64_t)A.ld, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W.hi, (uint64_t)N, (uint64_t)K, (uint64_t)W.ld, (uint32_t)(BN / 2), 64); if (rc) return rc;
  tmAl = tmA; tmBl = tmB;
  if (X3) {
    if (CONV) rc = make_tmap_bf16_nhwc(&tmAl, A.lo, (uint64_t)conv_b, (uint64_t)conv_h, (uint64_t)conv_w, (uint64_t)conv_cin);
    else rc = make_tmap_bf16_2d(&tmAl, A.lo, (uint64_t)M, (uint64_t)K, (uint64_t)A.ld, 128, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmBl, W.lo, (uint64_t)N, (uint64_t)K, (uint64_t)W.ld, (uint32_t)(BN / 2), 64); if (rc) return rc;
  }
  PairParams p;
  p.M = M; p.N = N; p.K = K;
  p.tiles_n = N / BN; p.splits = splits;
  const long long units = (long long)(M / 256) * p.tiles_n * splits;
  DVD_REQUIRE(units < (1LL << 31), "gemm_pair: too many work units");
  p.units = (int)units;
  p.npairs = units < npairs_max ? (int)units : npairs_max;
  p.conv_h = conv_h; p.conv_w = conv_w; p.conv_cin = conv_cin;
  p.partial = sk ? sk->partial : nullptr; p.counters = sk ? sk->counters : nullptr;
  DVD_CUDA(launch_pdl_cluster(1, kern, dim3(2 * p.npairs), dim3(PP_THREADS), (size_t)Cfg::SMEM, st, 2, 1, tmA, tmAl, tmB, tmBl, p, e));
  DVD_LAUNCH_CHECK("k_gemm_pair");
  return 0;
}

template <bool X3, bool CONV>
static int launch_pair_bn(int bn, const TcMat& A, const TcMat& W, int M, int N, int K, const Epilogue& e, int conv_b, int conv_h, int conv_w,
                          int conv_cin, int splits, const TcScratch* sk, cudaStream_t st) {
  switch (bn) {
    case 256: return launch_pair<256, X3, CONV>(A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, splits, sk, st);
    case 192: return launch_pair<192, X3, CONV>(A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, splits, sk, st);
    case 128: return launch_pair<128, X3, CONV>(A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, splits, sk, st);
    default:  return launch_pair<64, X3, CONV>(A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, splits, sk, st);
  }
}

// A: 2-D [M,K] or, when conv_h > 0, NHWC [B,H,W,Cin] with K = 9*Cin; W: [N,K] K-major.
int gemm_pair_dispatch(const TcMat& A, const TcMat& W, int M, int N, int K, const Epilogue& e, int conv_b, int conv_h, int conv_w, int conv_cin,
                       const TcScratch* sk, cudaStream_t st) {
  read_env();
  const bool conv = conv_h > 0, x3 = A.lo != nullptr;
  DVD_REQUIRE(gemm_pair_supported(M, N, K, conv), "gemm_pair: unsupported shape M=%d N=%d K=%d", M, N, K);
  DVD_REQUIRE(!conv || (conv_w % 128 == 0 && conv_cin % 64 == 0 && K == 9 * conv_cin), "gemm_pair: bad conv geometry");
  int bn = 64, splits = 1;
  const bool can_split = !conv && !e.vt_out;
  pick_config(M, N, K, x3, can_split, sk, sm_count() / 2, bn, splits);
  if (conv) return x3 ? launch_pair_bn<true, true>(bn, A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, 1, nullptr, st)
                      : launch_pair_bn<false, true>(bn, A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, 1, nullptr, st);
  return x3 ? launch_pair_bn<true, false>(bn, A, W, M, N, K, e, 0, 0, 0, 0, splits, sk, st)
            : launch_pair_bn<false, false>(bn, A, W, M, N, K, e, 0, 0, 0, 0, splits, sk, st);
}

}  // namespace dvd
