// Persistent CTA-pair tcgen05 GEMM: the kernel every dense layer of the denoiser runs on.
//
//   C[M,N] = epilogue(A[M,K] * W[N,K]^T)      A, W K-major 16-bit (bf16, or bf16 hi + lo pairs = 3 tensor-core passes, DVD_PREC_BF16X3)
//
// Shape of the machine: 74 clusters of two CTAs (the two SMs of a TPC), one cluster per TPC, each looping over work units.
//   * a unit = one 256 x BN output tile (UMMA M = 256 across the pair, tcgen05.mma.cta_group::2); units are dealt round-robin
//     (unit u -> pair u % npairs).  The M = 2048 problems of a single document have only 48..144 tiles: the host picks BN in
//     {64,128,192,256} from a cycle model so that the tile count fits the 74 pairs best (e.g. N = 1536 runs as 64 tiles of 256 x 192).
//     Split-K through global memory was implemented and measured (profiles/r2_gemm_sweep.txt): always slower - the partial tiles'
//     write / wait / read-back serialises two epilogues per tile - and was removed again.
//   * per SM the pair mode stages 128 A rows and only HALF of the W tile (BN/2 rows): 32 KB instead of 48 KB of L2 -> SM traffic per
//     k-block of a 128 x 256 tile, which is what bounds a one-SM-per-tile kernel (~13 TB/s of L2 reads at 1.1 PFLOP/s).
//   * warp 0 (one lane, both CTAs)  TMA producer: ring of 3..8 stages of [A_hi | A_lo | W_hi | W_lo] boxes (SWIZZLE_128B), continuous
//     across units; transaction bytes of both CTAs are counted on the LEADER's full barrier.
//   * warp 1 (one lane, leader)     MMA issuer: 4 (or 12: hi*hi, lo*hi, hi*lo) UMMA 256 x BN x 16 per stage into one of TWO TMEM
//     accumulators; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs.
//   * warps 2..9 (both CTAs)        epilogue of the CTA's own 128 rows, overlapping the next unit's main loop: each warp owns a TMEM
//     lane quarter x half of the columns; tcgen05.ld 32x32 -> private XOR-swizzled 4 KB staging tile -> row-coalesced fused epilogue
//     (8 lanes per 128-byte row segment).
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
  int tiles_n, units, npairs;
  int conv_h, conv_w, conv_cin;
  int epi;                        // epi_key(flags, out) of a compiled epilogue body, or EPI_GENERIC
};

// ---- cluster / pair PTX helpers
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
// arrive on a barrier of (possibly) the other CTA of the pair.  Default semantics (.release.cta): the orderings that matter here are
// carried by tcgen05.fence / TMA complete_tx, and a .release.cluster arrive costs a MEMBAR.ALL.GPU (~0.7 us, measured) every time.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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

// ---- optional phase trace (-DDVD_GEMM_TRACE, tools/gemm_trace.py): %globaltimer per CTA at the phase boundaries of its first two units
#ifdef DVD_GEMM_TRACE
__device__ unsigned long long g_pair_trace[512][16];
__device__ __forceinline__ unsigned long long ptime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PTRACE(slot) do { if ((slot) < 16) g_pair_trace[blockIdx.x & 511][slot] = ptime(); } while (0)
#else
#define PTRACE(slot) do { } while (0)
#endif
#ifdef DVD_GEMM_TRACE2      // epilogue detail of (unit 0, chunk 0, warp 2) in slots 7..12 (single-unit shapes only)
#define ETRACE(slot) do { if (warp == 2 && lane == 0 && it == 0 && ch == 0) PTRACE(slot); } while (0)
#else
#define ETRACE(slot) do { } while (0)
#endif

// ---- specialised epilogue bodies.  The generic Epilogue is a bag of run-time options; evaluating them per row costs ~600 warp
// instructions per 32 x 32 chunk, and with two epilogue warps per scheduler the epilogue of a 128 x 256 tile took 10 us (measured
// with tools/gemm_trace.py --epi).  The host classifies the Epilogue into a (flags, output kind) key; the combinations the denoiser
// uses are compiled as straight-line code (~150 instructions per chunk), anything else takes the generic path.
enum { EF_SCALE = 1, EF_FLOOR = 2, EF_GATE = 4, EF_POS = 8, EF_RES = 16, EF_GELU = 32, EF_GELUX = 64 };
enum { EO_F32 = 0, EO_BF16 = 1, EO_PAIR = 2, EO_F16 = 3 };
constexpr int EPI_GENERIC = -1;
__host__ __device__ constexpr int epi_key(int flags, int out) { return (flags << 2) | out; }

struct EpiRowCtx { int col, prow, res_row0, pos_row0, orow0, ocol_add, N; };

template <int F, int O>
__device__ __forceinline__ void epi_rows(const Epilogue& e, const float4 (&a)[8], const EpiRowCtx& c) {
  const float4 cb = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + c.col)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 cs = make_float4(1.f, 1.f, 1.f, 1.f), ct = make_float4(0.f, 0.f, 0.f, 0.f), cg = cs;
  if (F & EF_SCALE) { cs = __ldg(reinterpret_cast<const float4*>(e.scale + c.col)); ct = __ldg(reinterpret_cast<const float4*>(e.shift + c.col)); }
  if (F & EF_GATE) cg = __ldg(reinterpret_cast<const float4*>(e.gate + c.col));
  float4 qv[8], pv[8];
  if (F & EF_RES) {
    const float* rp = e.resid + (size_t)(c.res_row0 + c.prow) * e.ldr + c.col;       // may alias e.out (in-place residual)
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = *reinterpret_cast<const float4*>(rp + (size_t)(i * 4) * e.ldr);
  }
  if (F & EF_POS) {
    const float* pp = e.pos + (size_t)(c.pos_row0 + c.prow) * c.N + c.col;
#pragma unroll
    for (int i = 0; i < 8; ++i) pv[i] = __ldg(reinterpret_cast<const float4*>(pp + (size_t)(i * 4) * c.N));
  }
  const size_t o0 = (size_t)(c.orow0 + c.prow) * (O == EO_F32 ? e.ldc : e.ldc_bf16) + c.col + c.ocol_add;
  const size_t ostep = (size_t)4 * (O == EO_F32 ? e.ldc : e.ldc_bf16);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v[4] = {a[i].x + cb.x, a[i].y + cb.y, a[i].z + cb.z, a[i].w + cb.w};
    if (F & EF_SCALE) { v[0] = v[0] * cs.x + ct.x; v[1] = v[1] * cs.y + ct.y; v[2] = v[2] * cs.z + ct.z; v[3] = v[3] * cs.w + ct.w; }
    if (F & EF_FLOOR) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
    if (F & EF_GELU) { v[0] = gelu_tanh_fast(v[0]); v[1] = gelu_tanh_fast(v[1]); v[2] = gelu_tanh_fast(v[2]); v[3] = gelu_tanh_fast(v[3]); }
    if (F & EF_GELUX) { v[0] = gelu_tanh(v[0]); v[1] = gelu_tanh(v[1]); v[2] = gelu_tanh(v[2]); v[3] = gelu_tanh(v[3]); }
    if (F & EF_POS) { v[0] += pv[i].x; v[1] += pv[i].y; v[2] += pv[i].z; v[3] += pv[i].w; }
    if (F & EF_GATE) { v[0] *= cg.x; v[1] *= cg.y; v[2] *= cg.z; v[3] *= cg.w; }
    if (F & EF_RES) { v[0] += qv[i].x; v[1] += qv[i].y; v[2] += qv[i].z; v[3] += qv[i].w; }
    const size_t off = o0 + i * ostep;
    if (O == EO_F32) {
      *reinterpret_cast<float4*>(e.out + off) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (O == EO_PAIR) {
      uint2 uu, ll;
      split_bf16x2(v[0], v[1], uu.x, ll.x);
      split_bf16x2(v[2], v[3], uu.y, ll.y);
      *reinterpret_cast<uint2*>(e.out_bf16 + off) = uu;
      *reinterpret_cast<uint2*>(e.out_lo + off) = ll;
    } else if (O == EO_F16) {
      *reinterpret_cast<uint2*>(e.out_bf16 + off) = make_uint2(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]));
    } else {
      *reinterpret_cast<uint2*>(e.out_bf16 + off) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
  }
}

// any other Epilogue: run-time options (kept out of line: it is not on the denoiser's path)
__device__ __noinline__ void epi_rows_generic(const Epilogue& e, const float4 (&a)[8], const EpiRowCtx& c) {
  const EpiCols ec = load_epi_cols(e, c.col);
  const bool has_res = e.resid != nullptr, has_pos = e.pos != nullptr;
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), pv = qv;
    if (has_res) qv = *reinterpret_cast<const float4*>(e.resid + (size_t)(c.res_row0 + c.prow + 4 * i) * e.ldr + c.col);
    if (has_pos) pv = __ldg(reinterpret_cast<const float4*>(e.pos + (size_t)(c.pos_row0 + c.prow + 4 * i) * c.N + c.col));
    float v[4];
    apply_epi4(ec, a[i], has_pos, pv, has_res, qv, v);
    store_tc_out4(e, c.orow0 + c.prow + 4 * i, c.col + c.ocol_add, v);
  }
}

// classification of an Epilogue (host): EPI_GENERIC when it is not one of the compiled combinations
static int classify_epilogue(const Epilogue& e) {
  int f = 0;
  if (e.scale) f |= EF_SCALE;
  if (e.act == ACT_RELU) f |= EF_FLOOR;
  else if (e.act == ACT_GELU) f |= EF_GELU;
  else if (e.act == ACT_GELU_EXACT) f |= EF_GELUX;
  else if (e.act != ACT_NONE) return EPI_GENERIC;
  if (e.gate) f |= EF_GATE;
  if (e.pos) f |= EF_POS;
  if (e.resid) f |= EF_RES;
  int o;
  if (e.out && !e.out_bf16) o = EO_F32;
  else if (!e.out && e.out_bf16) o = e.out_lo ? EO_PAIR : (e.out_f16 ? EO_F16 : EO_BF16);
  else return EPI_GENERIC;
  const int k = epi_key(f, o);
  switch (k) {
    case epi_key(0, EO_F32): case epi_key(0, EO_BF16): case epi_key(0, EO_PAIR): case epi_key(0, EO_F16):
    case epi_key(EF_POS, EO_BF16): case epi_key(EF_POS, EO_PAIR):
    case epi_key(EF_RES, EO_F32): case epi_key(EF_GATE | EF_RES, EO_F32):
    case epi_key(EF_GELU, EO_BF16): case epi_key(EF_GELUX, EO_PAIR):
    case epi_key(EF_SCALE | EF_FLOOR, EO_BF16): case epi_key(EF_SCALE | EF_FLOOR, EO_PAIR): case epi_key(EF_SCALE | EF_FLOOR | EF_RES, EO_F32):
    case epi_key(EF_FLOOR, EO_BF16): case epi_key(EF_FLOOR, EO_PAIR):
      return k;
    default:
      return EPI_GENERIC;
  }
}

struct PairUnit { int m0, n0, kb0, kb1; };
__device__ __forceinline__ PairUnit decode_unit(const PairParams& p, int u, int nkb, int bn) {
  PairUnit q;
  const int mt = u / p.tiles_n, nt = u - mt * p.tiles_n;
  q.m0 = mt * (2 * PBM); q.n0 = nt * bn;
  q.kb0 = 0; q.kb1 = nkb;
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
    PTRACE(0);                                               // CTA start
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    if (X3) { prefetch_tmap(&tmAl); prefetch_tmap(&tmBl); }
    for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }       // full: the leader's arrive.expect_tx; both CTAs' TMA bytes
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 2 * PP_EPI_WARPS); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
  fence_before_sync();
  cluster_sync_all();                                        // the barriers of BOTH CTAs exist before any remote arrive / TMA
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) PTRACE(1);                           // prologue done (the dependency wait follows per role)

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own 128 A rows, own half of the W tile; the bytes of both CTAs are counted on the LEADER's full
      // barrier, which only the leader arrives on (expect_tx of both CTAs' bytes): a follower-side arrive per k-block would put a
      // cluster-scope fence on the producer's critical path.  The follower refills a stage only after ITS empty barrier fired, i.e. after
      // the previous phase of the leader's full barrier completed, so bytes never land in the wrong phase.
      // The W tiles do not depend on the previous kernel: the first ring-full of them is requested BEFORE griddepcontrol.wait, so that
      // under programmatic dependent launch the weight fetch (HBM: a step streams more weights than the L2 holds) overlaps the
      // predecessor's tail; the A tiles follow after the wait.
      uint32_t g = 0;
      const int cblocks = CONV ? p.conv_cin / 64 : 1;
      bool waited_pdl = false;
      for (int u = pair; u < p.units; u += p.npairs) {
        const PairUnit q = decode_unit(p, u, nkb, BN);
        const int m0 = q.m0 + (int)rank * PBM, nb = q.n0 + (int)rank * (BN / 2);
        int cn = 0, cy = 0, cx = 0;
        if (CONV) {
          const int hw = p.conv_h * p.conv_w;
          cn = m0 / hw; const int rem = m0 - cn * hw; cy = rem / p.conv_w; cx = rem - cy * p.conv_w;   // 128 consecutive pixels of one image row
        }
        auto load_b = [&](int kb, uint32_t gg) {
          const int s = gg % ST;
          uint8_t* a = smem + s * Cfg::STAGE_BYTES;
          const uint32_t lead_full = mapa(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);        // both CTAs' loads of this stage
#pragma unroll
          for (int o = 0; o < Cfg::NOP; ++o) tma_load_2d_pair(a + Cfg::B_OFF + o * Cfg::B_BYTES, o ? &tmBl : &tmB, lead_full, kb * PBK, nb);
        };
        auto load_a = [&](int kb, uint32_t gg) {
          const int s = gg % ST;
          uint8_t* a = smem + s * Cfg::STAGE_BYTES;
          const uint32_t lead_full = mapa(smem_u32(&full[s]), 0);
#pragma unroll
          for (int o = 0; o < Cfg::NOP; ++o) {
            if (CONV) {
              const int tap = kb / cblocks, cb = kb - tap * cblocks;
              tma_load_4d_pair(a + o * Cfg::A_BYTES, o ? &tmAl : &tmA, lead_full, cb * 64, cx + tap % 3 - 1, cy + tap / 3 - 1, cn);
            } else {
              tma_load_2d_pair(a + o * Cfg::A_BYTES, o ? &tmAl : &tmA, lead_full, kb * PBK, m0);
            }
          }
        };
        int kb = q.kb0;
        if (!waited_pdl) {                                     // first unit: W tiles of the first ring-full ahead of the dependency wait
          const int npre = (q.kb1 - q.kb0) < ST ? (q.kb1 - q.kb0) : ST;
          for (int i = 0; i < npre; ++i) load_b(q.kb0 + i, g + i);
          pdl_wait();
          waited_pdl = true;
          for (int i = 0; i < npre; ++i) load_a(q.kb0 + i, g + i);
          kb += npre; g += npre;
        }
        for (; kb < q.kb1; ++kb, ++g) {
          mbar_wait(&empty[g % ST], ((g / ST) & 1) ^ 1);
          load_b(kb, g);
          load_a(kb, g);
        }
      }
      if (!waited_pdl) pdl_wait();
    } else {
      pdl_wait();
    }
    __syncwarp();
  } else if (warp == 1) {
    pdl_wait();
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
          if (kb == q.kb0) if (it < 2) PTRACE(2 + 5 * it);                // first stage of the unit landed
          if (kb == q.kb1 - 1) if (it < 2) PTRACE(3 + 5 * it);            // last stage landed
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
    pdl_wait();
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
      const int res_row0 = e.resid_mod ? rbase % e.resid_mod : rbase;
      const int pos_row0 = e.pos_rows ? rbase % e.pos_rows : 0;
      int orow0 = rbase, ocol_add = 0;
      if (e.group_rows) { orow0 = rbase % e.group_rows; ocol_add = (rbase / e.group_rows) * e.group_col_stride; }
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      if (warp == 2 && lane == 0 && it < 2) PTRACE(4 + 5 * it);          // accumulator complete
      fence_after_sync();
      const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * (BN / 2));
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        const int col0 = q.n0 + half * (BN / 2) + ch * 32;
        // ---- phase 1 (thread = row = TMEM lane): TMEM -> registers -> swizzled staging (+ transposed V^T store, coalesced in this mapping)
        uint32_t r[32];
        tmem_ld_32x32(tacc + (uint32_t)(ch * 32), r);
        tmem_ld_wait();
        ETRACE(7);
        if (ch == NCH - 1) {                                    // accumulator fully copied out: hand the buffer back to the MMA warp
          if (warp == 2 && lane == 0 && it < 2) PTRACE(5 + 5 * it);      // TMEM drained
          fence_before_sync();
          __syncwarp();
          // (nobody waits for the buffers of a pair's last two units)
          if (lane == 0 && u + 2 * p.npairs < p.units) mbar_arrive_cluster(lead_acc_empty0 + (uint32_t)buf * 8u);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        if (e.vt_out && col0 >= e.vt_col0) {                    // bias-only epilogue (host check)
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
        ETRACE(8);
        // ---- phase 2 (8 lanes = one 128-byte row segment): staging -> fused epilogue -> global
        const int col = col0 + 4 * pc;
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + prow;
          a[i] = *reinterpret_cast<const float4*>(stage + rr * 32 + ((pc ^ (rr & 7)) << 2));
        }
        ETRACE(9);
        {
          // Row bookkeeping happens once per unit (rows of a 32-row group never straddle a residual / pos-embed / stream boundary: the
          // host checks that those periods are multiples of 128), so the per-row work is address increments only.
          const EpiRowCtx rc{col, prow, res_row0, pos_row0, orow0, ocol_add, p.N};
          ETRACE(10);
#define DVD_EPI_CASE(F, O) case epi_key(F, O): epi_rows<F, O>(e, a, rc); break;
          switch (p.epi) {
            DVD_EPI_CASE(0, EO_F32) DVD_EPI_CASE(0, EO_BF16) DVD_EPI_CASE(0, EO_PAIR) DVD_EPI_CASE(0, EO_F16)
            DVD_EPI_CASE(EF_POS, EO_BF16) DVD_EPI_CASE(EF_POS, EO_PAIR)
            DVD_EPI_CASE(EF_RES, EO_F32) DVD_EPI_CASE(EF_GATE | EF_RES, EO_F32)
            DVD_EPI_CASE(EF_GELU, EO_BF16) DVD_EPI_CASE(EF_GELUX, EO_PAIR)
            DVD_EPI_CASE(EF_SCALE | EF_FLOOR, EO_BF16) DVD_EPI_CASE(EF_SCALE | EF_FLOOR, EO_PAIR) DVD_EPI_CASE(EF_SCALE | EF_FLOOR | EF_RES, EO_F32)
            DVD_EPI_CASE(EF_FLOOR, EO_BF16) DVD_EPI_CASE(EF_FLOOR, EO_PAIR)
            default: epi_rows_generic(e, a, rc); break;
          }
#undef DVD_EPI_CASE
        }
        ETRACE(11);
        __syncwarp();                                           // the staging tile is rewritten by the next chunk
        ETRACE(12);
      }
      if (warp == 2 && lane == 0 && it < 2) PTRACE(6 + 5 * it);          // epilogue of the unit done (this warp)
    }
  }
  if (threadIdx.x == 64) PTRACE(14);                          // this CTA's epilogue warps are done
  fence_before_sync();
  cluster_sync_all();                                          // the peer's smem / TMEM must outlive every MMA that reads it
  if (warp == 1) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
  if (threadIdx.x == 32) PTRACE(15);                          // exit
}

// ---------------------------------------------------------------------------------------- host side
static int g_force_bn = -1, g_debug = 0;
static void read_env() {
  if (g_force_bn >= 0) return;
  const char* b = getenv("DVD_GEMM_BN"); g_force_bn = b ? atoi(b) : 0;
  const char* d = getenv("DVD_GEMM_DEBUG"); g_debug = d ? atoi(d) : 0;
}

bool gemm_pair_supported(int M, int N, int K, bool conv) {
  (void)K; (void)conv;
  return M % 256 == 0 && N % 64 == 0;
}
// the epilogue computes residual / pos-embed / stream-concat row offsets once per 32-row group
static bool epilogue_periods_ok(const Epilogue& e) {
  return (e.resid_mod % 128 == 0) && (e.pos_rows % 128 == 0) && (e.group_rows % 128 == 0);
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

// Cycle model used to pick the tile width.  Per k-block and CTA: tensor time (UMMA 256 x bn x 16 on a pair = bn/2 clocks, 4 per k-block,
// x3 in split-precision mode) against the L2 -> SM operand stream (~43 B/clk per SM when every SM pulls: 12.5 TB/s measured with
// tools/gemm_trace.py).  Per unit a fixed pipeline-fill cost, per launch the exposed last epilogue.  Calibrated against
// profiles/r2_gemm_sweep.txt (all denoiser shapes x {128,192,256}).
static double unit_kb_cycles(int bn, bool x3) {
  const double mma = (x3 ? 3.0 : 1.0) * 4.0 * (bn / 2.0) * 1.15;
  const double bytes = (x3 ? 2.0 : 1.0) * (16384.0 + bn * 64.0);
  const double mem = bytes / 43.0;
  return mma > mem ? mma : mem;
}

static int pick_bn(int M, int N, int K, bool x3, int npairs) {
  const int nkb = (K + PBK - 1) / PBK;
  double best = -1.0;
  int bn_out = 64;
  const int bns[4] = {256, 192, 128, 64};
  for (int bi = 0; bi < 4; ++bi) {
    const int bn = bns[bi];
    if (N % bn) continue;
    if (g_force_bn > 0 && bn != g_force_bn && N % g_force_bn == 0) continue;
    const long long units = (long long)(M / 256) * (N / bn);
    const long long waves = (units + npairs - 1) / npairs;
    const double cost = waves * (nkb * unit_kb_cycles(bn, x3) + 700.0) + (900.0 + 9.0 * bn);
    if (best < 0 || cost < best) { best = cost; bn_out = bn; }
  }
  return bn_out;
}

template <int BN, bool X3, bool CONV>
static int launch_pair(const TcMat& A, const TcMat& W, int M, int N, int K, const Epilogue& e, int conv_b, int conv_h, int conv_w, int conv_cin,
                       cudaStream_t st) {
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
  p.tiles_n = N / BN;
  const long long units = (long long)(M / 256) * p.tiles_n;
  DVD_REQUIRE(units < (1LL << 31), "gemm_pair: too many work units");
  p.units = (int)units;
  p.npairs = units < npairs_max ? (int)units : npairs_max;
  p.conv_h = conv_h; p.conv_w = conv_w; p.conv_cin = conv_cin;
  p.epi = classify_epilogue(e);
  if (g_debug) fprintf(stderr, "[gemm_pair] M=%d N=%d K=%d x3=%d conv=%d bn=%d units=%d npairs=%d (max %d) stages=%d smem=%d\n", M, N, K,
                       (int)X3, (int)CONV, BN, p.units, p.npairs, npairs_max, Cfg::STAGES, Cfg::SMEM);
  DVD_CUDA(launch_pdl_cluster(1, kern, dim3(2 * p.npairs), dim3(PP_THREADS), (size_t)Cfg::SMEM, st, 2, 1, tmA, tmAl, tmB, tmBl, p, e));
  DVD_LAUNCH_CHECK("k_gemm_pair");
  return 0;
}

template <bool X3, bool CONV>
static int launch_pair_bn(int bn, const TcMat& A, const TcMat& W, int M, int N, int K, const Epilogue& e, int conv_b, int conv_h, int conv_w,
                          int conv_cin, cudaStream_t st) {
  switch (bn) {
    case 256: return launch_pair<256, X3, CONV>(A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, st);
    case 192: return launch_pair<192, X3, CONV>(A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, st);
    case 128: return launch_pair<128, X3, CONV>(A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, st);
    default:  return launch_pair<64, X3, CONV>(A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, st);
  }
}

// A: 2-D [M,K] or, when conv_h > 0, NHWC [B,H,W,Cin] with K = 9*Cin; W: [N,K] K-major.  W must not be written by the kernel that
// immediately precedes this launch in the stream if that kernel triggers programmatic dependent launch (the weight tiles are
// requested ahead of griddepcontrol.wait); the denoiser's weights are static.
int gemm_pair_dispatch(const TcMat& A, const TcMat& W, int M, int N, int K, const Epilogue& e, int conv_b, int conv_h, int conv_w, int conv_cin,
                       cudaStream_t st) {
  read_env();
  const bool conv = conv_h > 0, x3 = A.lo != nullptr;
  DVD_REQUIRE(gemm_pair_supported(M, N, K, conv), "gemm_pair: unsupported shape M=%d N=%d K=%d", M, N, K);
  DVD_REQUIRE(epilogue_periods_ok(e), "gemm_pair: resid_mod / pos_rows / group_rows must be multiples of 128");
  DVD_REQUIRE(!conv || (conv_w % 128 == 0 && conv_cin % 64 == 0 && K == 9 * conv_cin), "gemm_pair: bad conv geometry");
  const int bn = pick_bn(M, N, K, x3, sm_count() / 2);
  if (conv) return x3 ? launch_pair_bn<true, true>(bn, A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, st)
                      : launch_pair_bn<false, true>(bn, A, W, M, N, K, e, conv_b, conv_h, conv_w, conv_cin, st);
  return x3 ? launch_pair_bn<true, false>(bn, A, W, M, N, K, e, 0, 0, 0, 0, st)
            : launch_pair_bn<false, false>(bn, A, W, M, N, K, e, 0, 0, 0, 0, st);
}

}  // namespace dvd

#ifdef DVD_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) int dvd_debug_pair_trace(unsigned long long* out, int n_ctas) {
  return (int)cudaMemcpyFromSymbol(out, dvd::g_pair_trace, (size_t)n_ctas * 16 * sizeof(unsigned long long));
}
#endif
