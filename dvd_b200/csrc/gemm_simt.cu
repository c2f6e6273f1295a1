#include "gemm_simt.cuh"

namespace dvd {

constexpr int BM = 128, BK = 16, PAD = 4;

template <int AMODE, int BMODE, int BN>
__global__ void __launch_bounds__(256) k_gemm_f32(GemmParams p) {
  constexpr int TN = BN / 16;                       // columns per thread (8 or 4)
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
  const int z = blockIdx.z;
  const int zn = z / p.heads, zh = z % p.heads;
  const float* __restrict__ A = p.A + zn * p.sAn + zh * p.sAh;
  const float* __restrict__ B = p.B + (zn / p.bdiv) * p.sBn + zh * p.sBh;

  // ---- global -> register staging maps
  // A (and B in NK mode): row = tid/4 (+64), k quad = (tid%4)*4
  const int lrow = tid >> 2, lkq = (tid & 3) * 4;
  long long a_off[2]; bool a_ok[2];
  int cy[2], cx[2];                                 // conv: pixel coords
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int r = bm + lrow + h * 64;
    a_ok[h] = r < p.M;
    if (AMODE == A_DIRECT) {
      a_off[h] = (long long)r * p.lda;
    } else {
      int hw = p.convH * p.convW;
      int img = r / hw, rem = r % hw;
      cy[h] = rem / p.convW; cx[h] = rem % p.convW;
      a_off[h] = ((long long)img * hw) * p.convC;   // image base
    }
  }
  float4 ra[2], rb[2];

  auto load_a = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int k = k0 + lkq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[h] && k < p.K) {
        if (AMODE == A_DIRECT) {
          v = __ldg(reinterpret_cast<const float4*>(A + a_off[h] + k));
        } else {
          int tap = k / p.convC, c = k - tap * p.convC;
          int yy = cy[h] + tap / 3 - 1, xx = cx[h] + tap % 3 - 1;
          if (yy >= 0 && yy < p.convH && xx >= 0 && xx < p.convW)
            v = __ldg(reinterpret_cast<const float4*>(A + a_off[h] + ((long long)yy * p.convW + xx) * p.convC + c));
        }
      }
      ra[h] = v;
    }
  };
  auto load_b = [&](int k0) {
    if (BMODE == B_NK) {
#pragma unroll
      for (int h = 0; h < BN / 64; ++h) {
        int n = bn + lrow + h * 64, k = k0 + lkq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < p.N && k < p.K) v = __ldg(reinterpret_cast<const float4*>(B + (long long)n * p.ldb + k));
        rb[h] = v;
      }
    } else {
      constexpr int Q = BN / 4;                     // float4 per k-row
      constexpr int R = 256 / Q;                    // k-rows per pass
#pragma unroll
      for (int h = 0; h < BK / R; ++h) {
        int k = k0 + tid / Q + h * R, n = bn + (tid % Q) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < p.K && n < p.N) v = __ldg(reinterpret_cast<const float4*>(B + (long long)k * p.ldb + n));
        rb[h] = v;
      }
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lrow + h * 64;
      As[buf][lkq + 0][r] = ra[h].x; As[buf][lkq + 1][r] = ra[h].y;
      As[buf][lkq + 2][r] = ra[h].z; As[buf][lkq + 3][r] = ra[h].w;
    }
    if (BMODE == B_NK) {
#pragma unroll
      for (int h = 0; h < BN / 64; ++h) {
        int r = lrow + h * 64;
        Bs[buf][lkq + 0][r] = rb[h].x; Bs[buf][lkq + 1][r] = rb[h].y;
        Bs[buf][lkq + 2][r] = rb[h].z; Bs[buf][lkq + 3][r] = rb[h].w;
      }
    } else {
      constexpr int Q = BN / 4;
      constexpr int R = 256 / Q;
#pragma unroll
      for (int h = 0; h < BK / R; ++h)
        *reinterpret_cast<float4*>(&Bs[buf][tid / Q + h * R][(tid % Q) * 4]) = rb[h];
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nt = (p.K + BK - 1) / BK;
  load_a(0); load_b(0);
  store_smem(0);
  __syncthreads();
  for (int t = 0; t < nt; ++t) {
    const int cur = t & 1;
    if (t + 1 < nt) { load_a((t + 1) * BK); load_b((t + 1) * BK); }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      if (TN == 8) {
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][(BN / 2) + tx * 4]);
        b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < nt) store_smem(cur ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  const Epilogue& e = p.e;
  float* __restrict__ Cb = e.out ? e.out + zn * p.sCn + zh * p.sCh : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int row = bm + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= p.M) continue;
#pragma unroll
    for (int jb = 0; jb < TN / 4; ++jb) {
      int col = bn + jb * (BN / 2) + tx * 4;
      if (col >= p.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (col + j < p.N) ? apply_epilogue(e, acc[i][jb * 4 + j] * p.alpha, row, col + j, p.N) : 0.f;
      int orow, ocol;
      epilogue_dest(e, row, col, orow, ocol);
      if (e.out) {
        float* o = Cb + (size_t)orow * e.ldc + ocol;
        if (col + 3 < p.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
          for (int j = 0; j < 4 && col + j < p.N; ++j) o[j] = v[j];
        }
      }
      if (e.out_bf16) {
        __nv_bfloat16* ob = e.out_bf16 + (size_t)orow * e.ldc_bf16 + ocol;
        for (int j = 0; j < 4 && col + j < p.N; ++j) ob[j] = __float2bfloat16_rn(v[j]);
      }
    }
  }
}

int gemm_f32(const GemmParams& p, int amode, int bmode, int batch, cudaStream_t st) {
  DVD_REQUIRE(p.A && p.B && (p.e.out || p.e.out_bf16), "gemm_f32: null pointer");
  DVD_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0 && (p.K % 4) == 0, "gemm_f32: bad shape M=%d N=%d K=%d", p.M, p.N, p.K);
  DVD_REQUIRE((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.B) & 15) == 0, "gemm_f32: operands must be 16B aligned");
  if (amode == A_DIRECT) DVD_REQUIRE(p.lda % 4 == 0, "gemm_f32: lda %% 4");
  else DVD_REQUIRE(p.convC % 4 == 0 && p.K == 9 * p.convC, "gemm_f32: conv needs C %% 4 == 0 and K == 9C");
  DVD_REQUIRE(p.ldb % 4 == 0, "gemm_f32: ldb %% 4");
  const bool narrow = (p.N <= 64);
  DVD_REQUIRE(cdiv(p.M, BM) <= 65535, "gemm_f32: M = %d needs more than 65535 row tiles (the fp32 parity mode handles up to 31 documents per call)", p.M);
  dim3 grid(cdiv(p.N, narrow ? 64 : 128), cdiv(p.M, BM), batch);
  DVD_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm_f32: grid too large");
#define DVD_GEMM_CASE(AM, BMd)                                                         \
  if (amode == AM && bmode == BMd) {                                                   \
    if (narrow) k_gemm_f32<AM, BMd, 64><<<grid, 256, 0, st>>>(p);                      \
    else        k_gemm_f32<AM, BMd, 128><<<grid, 256, 0, st>>>(p);                     \
  }
  DVD_GEMM_CASE(A_DIRECT, B_NK)
  DVD_GEMM_CASE(A_DIRECT, B_KN)
  DVD_GEMM_CASE(A_CONV3, B_NK)
#undef DVD_GEMM_CASE
  DVD_LAUNCH_CHECK("k_gemm_f32");
  return 0;
}

}  // namespace dvd
