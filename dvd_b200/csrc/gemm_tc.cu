// tcgen05 GEMM: C[M,N] = epilogue(A[M,K] * W[N,K]^T), bf16 operands (both K-major, exactly the torch Linear layouts),
// fp32 accumulation in tensor memory.
//
//   * one CTA computes a 128 x BN output tile (UMMA M=128, N=BN, K=16; BN in {128, 256}); 4 warps:
//       warp 0 / one lane : TMA producer  — cp.async.bulk.tensor 2-D boxes (128 x 64 bf16, SWIZZLE_128B) into a ring of stages
//       warp 1 / one lane : MMA issuer    — 4 x tcgen05.mma per stage, tcgen05.commit frees the stage / signals the epilogue
//       warps 0..3        : epilogue      — tcgen05.ld (32 lanes x 32 columns per warp), fused Epilogue, vector stores
//   * smem ring sized so that two CTAs share an SM (2 x <=112 KB, 2 x <=256 TMEM columns): the epilogue of one CTA overlaps the
//     main loop of the other, which replaces a persistent scheduler for these small (M = 2048..8192) problems.
//   * K tails (K = 1032 for the r-embedder) are zero-filled by TMA; M/N tails are masked in the epilogue.
#include "gemm_tc.cuh"
#include "tc_common.cuh"
#include <stdlib.h>
#include <algorithm>

namespace dvd {

using namespace tc;

// ---------------------------------------------------------------------------------------- tensor-map encode (driver entry point at run time)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DVD_E_NOTMA; }
  DVD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, "tensor map: base/stride must be 16-byte aligned (ld=%llu)",
              (unsigned long long)ld);
  DVD_REQUIRE(box_cols * 2 == 128 && box_rows <= 256, "tensor map: box must be 128 bytes wide");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r, (unsigned long long)rows,
                                     (unsigned long long)cols, (unsigned long long)ld); return DVD_E_NOTMA; }
  return 0;
}

// 4-D bf16 NHWC activation [N, H, W, C]; box = 64 channels x 128 pixels of one image row (implicit-GEMM conv A operand).
// Out-of-bounds coordinates (x = -1 / W, y = -1 / H: the conv's zero padding) are zero-filled by TMA.
int make_tmap_bf16_nhwc(CUtensorMap* out, const void* base, uint64_t n, uint64_t h, uint64_t w, uint64_t c) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DVD_E_NOTMA; }
  DVD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && c % 64 == 0 && w % 128 == 0, "nhwc tensor map: need C %% 64 == 0 and W %% 128 == 0");
  cuuint64_t gdim[4] = {c, w, h, n};
  cuuint64_t gstr[3] = {c * 2, w * c * 2, h * w * c * 2};
  cuuint32_t box[4] = {64, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (nhwc) failed (%d)", (int)r); return DVD_E_NOTMA; }
  return 0;
}

// Plain (unswizzled) 3-D tile map over an image stack: dims {w, h, planes} of `elem_bytes`-sized elements (4 = fp32, 1 = uint8),
// box {box_w, box_h, box_p}.  Used by the unwarp kernel to stage source windows; OOB elements are zero-filled (= zeros padding).
int make_tmap_image3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t w, uint64_t h, uint64_t planes, uint32_t box_w,
                      uint32_t box_h, uint32_t box_p) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DVD_E_NOTMA; }
  DVD_REQUIRE(elem_bytes == 4 || elem_bytes == 1, "image tensor map: fp32 or uint8 only");
  DVD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (w * elem_bytes) % 16 == 0 && (box_w * elem_bytes) % 16 == 0 && box_w <= 256 &&
                  box_h <= 256 && box_p <= 256, "image tensor map: alignment / box limits");
  cuuint64_t gdim[3] = {w, h, planes};
  cuuint64_t gstr[2] = {w * elem_bytes, w * h * elem_bytes};
  cuuint32_t box[3] = {box_w, box_h, box_p};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (image) failed (%d) w=%llu h=%llu", (int)r, (unsigned long long)w,
                                     (unsigned long long)h); return DVD_E_NOTMA; }
  return 0;
}

// ---------------------------------------------------------------------------------------- cluster helpers (TMA multicast variant)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one TMA box fetched from L2 ONCE and written to the same shared-memory offset of every CTA in `mask` (each destination's mbarrier
// at the same offset receives the transaction bytes)
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
               : "memory");
}
// arrive (once all previously issued MMAs of this CTA have completed) on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// ---------------------------------------------------------------------------------------- optional phase trace (-DDVD_GEMM_TRACE)
#ifdef DVD_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[2048][8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define DVD_TRACE(slot) do { g_gemm_trace[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) & 2047][slot] = gtime(); } while (0)
#else
#define DVD_TRACE(slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------- kernel
constexpr int TBM = 128, TBK = 64;
constexpr int TC_THREADS = 256;          // warp 0: TMA producer, warp 1: MMA issuer, all 8 warps: epilogue

template <int BN>
struct TcCfg {
  static constexpr int STAGES = (BN == 64) ? 4 : ((BN <= 128) ? 3 : 2);   // <= 96 KB of ring -> two CTAs per SM
  static constexpr int TMEM_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);  // power of two >= BN
  static constexpr int A_BYTES = TBM * TBK * 2, B_BYTES = BN * TBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int CW = BN < 128 ? BN : 128;                          // columns per epilogue pass
  static constexpr int STAGE_LD = CW + 4;                                 // fp32 staging row stride (16-byte aligned, conflict-free)
  static constexpr int STAGING_BYTES = 4 * 32 * STAGE_LD * 4;             // 4 row groups x 32 rows
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static_assert(STAGING_BYTES <= RING_BYTES, "epilogue staging must fit in the (drained) operand ring");
  static constexpr int SMEM = RING_BYTES + 1024 /*align slack*/ + 128 /*barriers*/;
};

struct ConvGeom { int H, W, Cin; };      // CONV: A is an NHWC activation, K = 9 * Cin ordered [ky][kx][Cin]

// CLN x CLM > 1: the CTAs of a (CLN, CLM, 1) cluster share operand tiles through TMA multicast.  The CLN CTAs of a cluster row work on
// the same 128 rows of A, the CLM CTAs of a cluster column on the same BN rows of W: every CTA fetches 1/CLN of the A tile and 1/CLM of
// the W tile and multicasts them, so the L2 -> SM traffic of the main loop (what bounds the 128x128 kernel: ~64 FLOP per loaded byte)
// drops by CLN resp. CLM.  A stage may be refilled once every CTA that receives data from this one has consumed it, so the MMA
// issuer's commit arrives on the `empty` barrier of all CTAs of its cluster row and column (CLN + CLM - 1 arrivals per phase).
template <int BN, bool CONV, int CLN = 1, int CLM = 1>
__global__ void __launch_bounds__(TC_THREADS, 2) k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                 int M, int N, int K, Epilogue e, ConvGeom cg) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CL = CLN * CLM;
  static_assert(!CONV || CL == 1, "the implicit-GEMM conv path is not clustered");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // M tiles beyond the 65535 limit of gridDim.y (the 512x512 pyramid levels of >= 32 documents) are folded into gridDim.z
  const int m0 = (blockIdx.z * gridDim.y + blockIdx.y) * TBM, n0 = blockIdx.x * BN;
  if (m0 >= M) return;                                  // whole CTA (ragged last z-slice)
  const int nkb = (K + TBK - 1) / TBK;

  pdl_trigger();                                        // the next kernel of the stream may start its own prologue
  if (threadIdx.x == 0) {
    DVD_TRACE(0);                                       // CTA start
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CLN + CLM - 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                       // peers' barriers are initialised before anything is multicast to them
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // position in the cluster: rank = x + y * CLN (grid dims are multiples of the cluster dims, checked on the host)
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  const int cxp = crank % CLN, cyp = crank / CLN;
  uint16_t mask_row = 0, mask_col = 0;                  // CTAs sharing my A tile / my W tile
#pragma unroll
  for (int i = 0; i < CLN; ++i) mask_row |= (uint16_t)(1u << (cyp * CLN + i));
#pragma unroll
  for (int i = 0; i < CLM; ++i) mask_col |= (uint16_t)(1u << (i * CLN + cxp));
  pdl_wait();                                           // everything above overlapped the previous kernel's tail
  if (threadIdx.x == 0) DVD_TRACE(1);                   // prologue done, predecessor finished

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      int cn = 0, cy = 0, cx = 0, cblocks = 1;
      if (CONV) {
        const int hw = cg.H * cg.W;
        cn = m0 / hw; const int rem = m0 % hw; cy = rem / cg.W; cx = rem % cg.W;      // 128 consecutive pixels of one image row
        cblocks = cg.Cin / 64;
      }
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES, it = kb / STAGES;
        mbar_wait(&empty[s], (it & 1) ^ 1);
        uint8_t* a = smem + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        if (CL > 1) {
          // my slice of the A tile (rows cxp * 128/CLN ...) to the cluster row, my slice of the W tile to the cluster column;
          // the tensor-map boxes are 128/CLN and min(128, BN/CLM) rows tall
          constexpr int AR = TBM / CLN, BR = BN / CLM, BBOX = BR > 128 ? 128 : BR;
          tma_load_2d_mc(a + cxp * AR * TBK * 2, &tmA, &full[s], kb * TBK, m0 + cxp * AR, mask_row);
#pragma unroll
          for (int j = 0; j < BR / BBOX; ++j)
            tma_load_2d_mc(a + Cfg::A_BYTES + (cyp * BR + j * BBOX) * TBK * 2, &tmB, &full[s], kb * TBK, n0 + cyp * BR + j * BBOX, mask_col);
        } else {
          if (CONV) {
            const int tap = kb / cblocks, cb = kb % cblocks;
            tma_load_4d(a, &tmA, &full[s], cb * 64, cx + tap % 3 - 1, cy + tap / 3 - 1, cn);
          } else {
            tma_load_2d(a, &tmA, &full[s], kb * TBK, m0);
          }
          tma_load_2d(a + Cfg::A_BYTES, &tmB, &full[s], kb * TBK, n0);
          if (BN == 256) tma_load_2d(a + Cfg::A_BYTES + 128 * TBK * 2, &tmB, &full[s], kb * TBK, n0 + 128);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(TBM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES, it = kb / STAGES;
        mbar_wait(&full[s], it & 1);
        if (kb == 0) DVD_TRACE(2);                      // first stage landed
        if (kb == nkb - 1) DVD_TRACE(3);                // last stage landed
        fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES), b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < TBK / 16; ++k) {
          // advancing K inside the 128-byte swizzle atom = +32 bytes on the start address
          mma_f16_ss(tmem_base, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc, (kb | k) ? 1u : 0u);
        }
        if (CL > 1) mma_commit_mc(&empty[s], mask_row | mask_col);   // every CTA that sends me operands learns the stage is free
        else        mma_commit(&empty[s]);               // stage reusable once these MMAs have read it
      }
      mma_commit(tmem_full);                             // accumulator complete
    }
    __syncwarp();
  }

  // ===== epilogue.  tmem_full => every MMA has retired, so every TMA write has been consumed: the operand ring is dead
  // and is reused as an fp32 staging tile.  The epilogue is a third of a batch-1 launch (tools/gemm_trace.py), so all 8 warps take
  // part: warps w and w+4 own the same 32 rows (a warp can only read the TMEM lanes 32*(w%4)..+31) and split the columns in
  // phase 1 and the rows in phase 2, synchronised by a 64-thread named barrier per row group.
  //   Phase 1 (thread = row, TMEM lane): TMEM -> registers -> staging (+ the transposed V^T store, which is naturally
  //   coalesced in this mapping).  Phase 2 (lane = 4 consecutive columns): staging -> fused epilogue -> fully coalesced
  //   512-byte row segments in global memory.
  mbar_wait(tmem_full, 0);
  if (threadIdx.x == 0) DVD_TRACE(4);                   // accumulator complete
  fence_after_sync();
  constexpr int CW = Cfg::CW, SLD = Cfg::STAGE_LD;
  static_assert((CW / 2) % 32 == 0, "each half of an epilogue pass is a whole number of 32-column TMEM loads");
  const int wq = warp & 3, wh = warp >> 2;               // row group (TMEM lane quarter), column / row half
  float* stage = reinterpret_cast<float*>(smem) + wq * 32 * SLD;
  const int row_t = m0 + wq * 32 + lane;                 // phase-1 row of this thread (M % 128 == 0 is checked on the host)
  auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory"); };
#pragma unroll 1
  for (int pass = 0; pass < BN / CW; ++pass) {
#pragma unroll 1
    for (int c0 = wh * (CW / 2); c0 < (wh + 1) * (CW / 2); c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(pass * CW + c0), r);
      tmem_ld_wait();
      float* srow = stage + lane * SLD + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(srow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                                         __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      const int col0 = n0 + pass * CW + c0;
      if (e.vt_out && col0 >= e.vt_col0 && col0 + 31 < N) {
        // V^T (bias-only epilogue, checked on the host): the bias of the 32 columns is warp-uniform -> 8 vector loads up front
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
    pair_sync();                                         // both column halves of the row group are staged
    // ---- phase 2: this warp finishes rows wh*16 .. wh*16+15 of the group
    const int col = n0 + pass * CW + 4 * lane;
    if (4 * lane < CW && col < N) {
      float4 cb = make_float4(0.f, 0.f, 0.f, 0.f), cs = make_float4(1.f, 1.f, 1.f, 1.f), ct = cb, cgate = cs;
      if (e.bias) cb = __ldg(reinterpret_cast<const float4*>(e.bias + col));
      if (e.scale) { cs = __ldg(reinterpret_cast<const float4*>(e.scale + col)); ct = __ldg(reinterpret_cast<const float4*>(e.shift + col)); }
      if (e.gate) cgate = __ldg(reinterpret_cast<const float4*>(e.gate + col));
      const bool has_scale = e.scale != nullptr, has_gate = e.gate != nullptr;
      const int act = e.act;
#pragma unroll 1
      for (int r0 = wh * 16; r0 < wh * 16 + 16; r0 += 8) {
        float4 a[8], q[8], p[8];
        // batch the loads of 8 rows (staging, residual, pos-embed) before any dependent math or store
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(stage + (r0 + i) * SLD + 4 * lane);
        if (e.resid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = m0 + wq * 32 + r0 + i;
            const int rr = e.resid_mod ? (row % e.resid_mod) : row;
            q[i] = *reinterpret_cast<const float4*>(e.resid + (size_t)rr * e.ldr + col);      // may alias e.out (in-place residual)
          }
        }
        if (e.pos) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = m0 + wq * 32 + r0 + i;
            p[i] = __ldg(reinterpret_cast<const float4*>(e.pos + (size_t)(row % e.pos_rows) * N + col));
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = m0 + wq * 32 + r0 + i;
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
    pair_sync();                                         // the staging rows may be overwritten by the next pass
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) DVD_TRACE(5);                   // epilogue done
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (CL > 1) cluster_sync_all();                       // peers may still be arriving on my `empty` barriers
}

template <int BN, int CLN, int CLM>
static int launch_tc_cluster(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const Epilogue& e, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DVD_CUDA(cudaFuncSetAttribute(k_gemm_tc<BN, false, CLN, CLM>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(N, BN), cdiv(M, TBM));
  ConvGeom cg{0, 0, 0};
  DVD_CUDA(launch_pdl_cluster(1, k_gemm_tc<BN, false, CLN, CLM>, grid, dim3(TC_THREADS), (size_t)TcCfg<BN>::SMEM, st, CLN, CLM, tmA, tmB, M, N, K, e, cg));
  DVD_LAUNCH_CHECK("k_gemm_tc (cluster)");
  return 0;
}

template <int BN, bool CONV>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const Epilogue& e, ConvGeom cg, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DVD_CUDA(cudaFuncSetAttribute(k_gemm_tc<BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM));
    attr_set = true;
  }
  const int my = cdiv(M, TBM), gy = my > 32768 ? 32768 : my;
  dim3 grid(cdiv(N, BN), gy, cdiv(my, gy));
  DVD_CUDA(launch_pdl(1, k_gemm_tc<BN, CONV>, grid, dim3(TC_THREADS), (size_t)TcCfg<BN>::SMEM, st, tmA, tmB, M, N, K, e, cg));
  DVD_LAUNCH_CHECK("k_gemm_tc");
  return 0;
}

static int check_epilogue(const Epilogue& e, int N) {
  DVD_REQUIRE(N % 4 == 0, "gemm_tc: N must be a multiple of 4");
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  DVD_REQUIRE(al(e.bias) && al(e.scale) && al(e.shift) && al(e.gate) && al(e.pos) && al(e.resid) && al(e.out) &&
              (reinterpret_cast<uintptr_t>(e.out_bf16) & 7) == 0, "gemm_tc: epilogue pointers must be 16-byte aligned");
  DVD_REQUIRE(e.ldc % 4 == 0 && e.ldr % 4 == 0 && e.ldc_bf16 % 4 == 0 && e.group_col_stride % 4 == 0, "gemm_tc: epilogue leading dims %% 4");
  DVD_REQUIRE(!e.vt_out || (e.vt_col0 % 32 == 0 && N % 32 == 0 && !e.scale && !e.act && !e.pos && !e.gate && !e.resid),
              "gemm_tc: the V^T output supports a bias-only epilogue with vt_col0 %% 32 == 0");
  return 0;
}

int gemm_tc2_dispatch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K, const Epilogue& e,
                      int conv_b, int conv_h, int conv_w, int conv_cin, cudaStream_t st);
static bool use_v1() {
  static int v = -1;
  if (v < 0) { const char* s = getenv("DVD_GEMM_V1"); v = (s && s[0] == '1') ? 1 : 0; }
  return v == 1;
}

int gemm_tc3_dispatch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K, const Epilogue& e, int bn,
                      cudaStream_t st);
static int v3_mode() {      // DVD_GEMM_V3: 0 = off, 1 = pair kernel wherever the shape allows, 2 = heuristic
  static int v = -1;
  if (v < 0) { const char* s = getenv("DVD_GEMM_V3"); v = s ? atoi(s) : 0; }
  return v;
}
static bool force_v2() {
  static int v = -1;
  if (v < 0) { const char* s = getenv("DVD_GEMM_V2"); v = (s && s[0] == '1') ? 1 : 0; }
  return v == 1;
}

int gemm_tc_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K, const Epilogue& e, cudaStream_t st) {
  DVD_REQUIRE(A && W && (e.out || e.out_bf16), "gemm_tc: null pointer");
  DVD_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && M % 128 == 0, "gemm_tc: bad shape M=%d N=%d K=%d (M must be a multiple of 128)", M, N, K);
  int rc = check_epilogue(e, N); if (rc) return rc;
  if (v3_mode() == 1 && M % 256 == 0 && N % 128 == 0) {
    int bn = (N % 256 == 0) ? 256 : 128;
    if (const char* f = getenv("DVD_GEMM_BN")) { int v = atoi(f); if ((v == 128 || v == 256) && N % v == 0) bn = v; }
    return gemm_tc3_dispatch(A, lda, W, ldw, M, N, K, e, bn, st);
  }
  // Measured on B200 (profiles/r1_gemm_microbench.txt): the persistent 128x256 kernel wins once there are >= 4 waves of wide
  // tiles (1.1 PFLOP/s at M = 16384); the small M = 2048 problems of a single document are latency-bound and run faster as two
  // co-resident 128x128 CTAs per SM.
  if (!use_v1() && (force_v2() || (N % 256 == 0 && (long long)(M / 128) * (N / 256) >= 4 * kSMs)))
    return gemm_tc2_dispatch(A, lda, W, ldw, M, N, K, e, 0, 0, 0, 0, st);
  // wide tiles only when they still fill the machine (>= ~1 wave of 2 CTAs/SM)
  bool wide = (N % 256 == 0) && ((long long)(M / 128) * (N / 256) * 100 >= 190LL * kSMs);   // >= ~1.9 SM-fulls of 128x256 tiles (2 CTAs/SM)
  if (const char* f = getenv("DVD_GEMM_WIDE")) wide = (N % 256 == 0) && atoi(f) == 1;     // tuning override
  const bool narrow = (N <= 64);
  // 128x96 and 128x192 (one CTA per SM, 5-stage ring) tiles were measured on the N = 1536 shapes: no gain over 128x128 (the main loop
  // already runs at the tensor rate of two co-resident CTAs; the launches are bound by prologue + epilogue, see tools/gemm_trace.py).
  CUtensorMap tmA, tmB;
  // TMA-multicast clusters (experiment, OFF by default): DVD_GEMM_CLUSTER=22 / 12 / 21 shares the A tile along N and / or the W tile
  // along M inside (2,2) / (1,2) / (2,1) clusters.  Measured on B200 (tools/gemm_bench.py): correct, but 5-10% SLOWER than independent
  // CTAs on every denoiser shape (e.g. 2048x1536x1536: 25.6 vs 23.6 us; 16384x4608x1536: 292 vs 235 us) - halving the L2 reads does not
  // help because the main loop is bound by what each SM can take in, and the cluster couples the progress of its CTAs.
  if (!narrow) {
    static const int cl_env = getenv("DVD_GEMM_CLUSTER") ? atoi(getenv("DVD_GEMM_CLUSTER")) : 0;
    const int bn = wide ? 256 : 128;
    const int gx = N / bn, gy = M / 128;
    int cl = 0;
    if (N % bn == 0 && cl_env != 0 && gy <= 32768) {
      if (gx % 2 == 0 && gy % 2 == 0) cl = 22; else if (gy % 2 == 0) cl = 12; else if (gx % 2 == 0) cl = 21;
      if (cl_env == 12 && gy % 2 == 0) cl = 12;
      if (cl_env == 21 && gx % 2 == 0) cl = 21;
    }
    if (cl) {
      const int cln = cl / 10, clm = cl % 10;
      rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128 / cln, 64); if (rc) return rc;
      rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, std::min(128, bn / clm), 64); if (rc) return rc;
      if (wide) {
        if (cl == 22) return launch_tc_cluster<256, 2, 2>(tmA, tmB, M, N, K, e, st);
        if (cl == 12) return launch_tc_cluster<256, 1, 2>(tmA, tmB, M, N, K, e, st);
        return launch_tc_cluster<256, 2, 1>(tmA, tmB, M, N, K, e, st);
      }
      if (cl == 22) return launch_tc_cluster<128, 2, 2>(tmA, tmB, M, N, K, e, st);
      if (cl == 12) return launch_tc_cluster<128, 1, 2>(tmA, tmB, M, N, K, e, st);
      return launch_tc_cluster<128, 2, 1>(tmA, tmB, M, N, K, e, st);
    }
  }
  rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64); if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, narrow ? 64 : 128, 64); if (rc) return rc;
  ConvGeom cg{0, 0, 0};
  if (wide) return launch_tc<256, false>(tmA, tmB, M, N, K, e, cg, st);
  if (narrow) return launch_tc<64, false>(tmA, tmB, M, N, K, e, cg, st);
  return launch_tc<128, false>(tmA, tmB, M, N, K, e, cg, st);
}

// 3x3 / pad 1 / stride 1 convolution as an implicit GEMM: in NHWC bf16 [B,H,W,Cin], Wt [Cout, 9*Cin] ([ky][kx][Cin]),
// out NHWC [B,H,W,Cout] through the Epilogue (bias + ReLU, bf16 and/or fp32).
int conv3x3_tc_bf16(const __nv_bfloat16* in, const __nv_bfloat16* Wt, int B, int H, int Wd, int Cin, int Cout, const Epilogue& e,
                    cudaStream_t st) {
  DVD_REQUIRE(in && Wt && (e.out || e.out_bf16), "conv3x3_tc: null pointer");
  DVD_REQUIRE(Cin % 64 == 0 && Wd % 128 == 0 && (Cout == 64 || Cout % 128 == 0), "conv3x3_tc: unsupported shape Cin=%d W=%d Cout=%d", Cin, Wd, Cout);
  const int M = B * H * Wd, K = 9 * Cin;
  int rc = check_epilogue(e, Cout); if (rc) return rc;
  if (!use_v1() && force_v2()) return gemm_tc2_dispatch(in, 0, Wt, K, M, Cout, K, e, B, H, Wd, Cin, st);
  CUtensorMap tmA, tmB;
  rc = make_tmap_bf16_nhwc(&tmA, in, (uint64_t)B, (uint64_t)H, (uint64_t)Wd, (uint64_t)Cin); if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, Wt, (uint64_t)Cout, (uint64_t)K, (uint64_t)K, Cout == 64 ? 64 : 128, 64); if (rc) return rc;
  ConvGeom cg{H, Wd, Cin};
  if (Cout == 64) return launch_tc<64, true>(tmA, tmB, M, Cout, K, e, cg, st);
  if (Cout % 256 == 0 && (long long)(M / 128) * (Cout / 256) >= 2 * kSMs) return launch_tc<256, true>(tmA, tmB, M, Cout, K, e, cg, st);
  return launch_tc<128, true>(tmA, tmB, M, Cout, K, e, cg, st);
}

}  // namespace dvd

#ifdef DVD_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) int dvd_debug_gemm_trace(unsigned long long* out, int n_ctas) {
  return (int)cudaMemcpyFromSymbol(out, dvd::g_gemm_trace, (size_t)n_ctas * 8 * sizeof(unsigned long long));
}
#endif
