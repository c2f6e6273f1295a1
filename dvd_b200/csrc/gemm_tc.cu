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

// ---------------------------------------------------------------------------------------- kernel
constexpr int TBM = 128, TBK = 64;

template <int BN>
struct TcCfg {
  static constexpr int STAGES = (BN == 128) ? 3 : 2;                    // 3 x 32 KB or 2 x 48 KB -> two CTAs per SM
  static constexpr int A_BYTES = TBM * TBK * 2, B_BYTES = BN * TBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 128 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(128) k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                 int M, int N, int K, Epilogue e) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
  const int nkb = (K + TBK - 1) / TBK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);            // BN fp32 accumulator columns (power of two >= 32)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES, it = kb / STAGES;
        mbar_wait(&empty[s], (it & 1) ^ 1);
        uint8_t* a = smem + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        tma_load_2d(a, &tmA, &full[s], kb * TBK, m0);
        tma_load_2d(a + Cfg::A_BYTES, &tmB, &full[s], kb * TBK, n0);
        if (BN == 256) tma_load_2d(a + Cfg::A_BYTES + 128 * TBK * 2, &tmB, &full[s], kb * TBK, n0 + 128);
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
        fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES), b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < TBK / 16; ++k) {
          // advancing K inside the 128-byte swizzle atom = +32 bytes on the start address
          mma_f16_ss(tmem_base, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc, (kb | k) ? 1u : 0u);
        }
        mma_commit(&empty[s]);                           // stage reusable once these MMAs have read it
      }
      mma_commit(tmem_full);                             // accumulator complete
    }
    __syncwarp();
  }

  // ===== epilogue: warp w owns TMEM lanes [32w, 32w+32) = output rows m0 + 32w + lane
  mbar_wait(tmem_full, 0);
  fence_after_sync();
  const int row = m0 + warp * 32 + lane;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tmem_ld_wait();
    const int col0 = n0 + c0;
    if (row < M && col0 < N) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (col0 + j < N) ? apply_epilogue(e, __uint_as_float(r[j]), row, col0 + j, N) : 0.f;
      int orow, ocol;
      epilogue_dest(e, row, col0, orow, ocol);
      const bool fullc = (col0 + 31 < N);
      if (e.out) {
        float* o = e.out + (size_t)orow * e.ldc + ocol;
        if (fullc && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 32 && col0 + j < N; ++j) o[j] = v[j];
        }
      }
      if (e.out_bf16) {
        __nv_bfloat16* o = e.out_bf16 + (size_t)orow * e.ldc_bf16 + ocol;
        if (fullc && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
            u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(o + j) = u;
          }
        } else {
          for (int j = 0; j < 32 && col0 + j < N; ++j) o[j] = __float2bfloat16_rn(v[j]);
        }
      }
      if (e.vt_out && col0 >= e.vt_col0) {
        // V^T: for a fixed column the 32 lanes hold 32 consecutive tokens -> 64-byte contiguous stores
        __nv_bfloat16* o = e.vt_out + ((size_t)(row >> 10) * (N - e.vt_col0) + (col0 - e.vt_col0)) * 1024 + (row & 1023);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) o[(size_t)j * 1024] = __float2bfloat16_rn(v[j]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

int gemm_tc_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K, const Epilogue& e, cudaStream_t st) {
  DVD_REQUIRE(A && W && (e.out || e.out_bf16), "gemm_tc: null pointer");
  DVD_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0, "gemm_tc: bad shape M=%d N=%d K=%d", M, N, K);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 128, 64);
  if (rc) return rc;
  // wide tiles only when they still fill the machine (>= ~1 wave of 2 CTAs/SM)
  const bool wide = (N % 256 == 0) && ((long long)(M / 128) * (N / 256) >= 2 * kSMs);
  static bool attr_set = false;
  if (!attr_set) {
    DVD_CUDA(cudaFuncSetAttribute(k_gemm_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM));
    DVD_CUDA(cudaFuncSetAttribute(k_gemm_tc<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256>::SMEM));
    attr_set = true;
  }
  if (wide) {
    dim3 grid(cdiv(N, 256), cdiv(M, TBM));
    k_gemm_tc<256><<<grid, 128, TcCfg<256>::SMEM, st>>>(tmA, tmB, M, N, K, e);
  } else {
    dim3 grid(cdiv(N, 128), cdiv(M, TBM));
    k_gemm_tc<128><<<grid, 128, TcCfg<128>::SMEM, st>>>(tmA, tmB, M, N, K, e);
  }
  DVD_LAUNCH_CHECK("k_gemm_tc");
  return 0;
}

}  // namespace dvd
