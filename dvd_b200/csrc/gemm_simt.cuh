// fp32 FFMA GEMM (DVD_PREC_FP32 mode) with fused epilogues; also the implicit-GEMM 3x3 conv of the
// pyramid and the batched QK^T / PV products of the fp32 attention path.
//
//   C[M,N] = epilogue( alpha * A[M,K] * B )      B given K-major (W[n][k], torch Linear layout)
//                                                 or N-major (B[k][n], used for P*V)
// 128 x BN x 16 tiles, 256 threads, 8 x (BN/16) register micro-tile, double-buffered smem with
// register-staged global prefetch.  This is the bit-reproducible reference mode of the library;
// the tensor-core (tcgen05) path in gemm_tc.cu shares the Epilogue description below.
#pragma once
#include "common.cuh"

namespace dvd {

// ACT_GELU: tanh-GELU; the tensor path evaluates it with the hardware tanh.approx (2^-11) unless ACT_GELU_EXACT (tanhf) is asked for
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SIGMOID = 3, ACT_GELU_EXACT = 4 };

struct Epilogue {
  const float* bias = nullptr;        // [N]
  const float* scale = nullptr;       // [N]  folded BN: v = v*scale + shift
  const float* shift = nullptr;
  int act = ACT_NONE;
  const float* pos = nullptr;         // [pos_rows, N] added by (row % pos_rows)
  int pos_rows = 0;
  const float* gate = nullptr;        // [N]  v *= gate
  const float* resid = nullptr;       // [resid_rows or M, ldr] residual added last
  int ldr = 0;
  int resid_mod = 0;                  // 0: resid row = row ; else row % resid_mod
  float* out = nullptr;
  int ldc = 0;
  int group_rows = 0;                 // 0: identity ; else out row = row % group_rows,
  int group_col_stride = 0;           //               out col += (row / group_rows) * group_col_stride
  __nv_bfloat16* out_bf16 = nullptr;  // optional 16-bit copy of the output (same mapping, ld = ldc_bf16): bf16, or fp16 if out_f16
  int ldc_bf16 = 0;
  // tensor path, split-precision mode (DVD_PREC_BF16X3): the operand of the NEXT GEMM is stored as a bf16 pair hi + lo with
  // hi = bf16(v), lo = bf16(v - hi) (same mapping and ld as out_bf16); operands of the fp16 attention kernel are stored as fp16.
  __nv_bfloat16* out_lo = nullptr;
  int out_f16 = 0;                    // out_bf16 (and vt_out) hold IEEE fp16 instead of bf16
  // tensor path only: columns >= vt_col0 (the V projection) are ALSO written transposed for the attention kernel:
  //   vt_out[(row / 1024) * (N - vt_col0) + (col - vt_col0)][row % 1024]      i.e. V^T [sample, C_v, T = 1024]
  __nv_bfloat16* vt_out = nullptr;
  int vt_col0 = 0;
  // tensor path (persistent pair kernel), LayerNorm fused into the CONSUMER GEMM: A holds the raw rows (16-bit operand), W is
  // pre-multiplied by the LN gain, and the epilogue applies
  //     v = rstd[row] * (acc - mean[row] * ln_colsum[col]) + bias[col]            (bias = W beta)
  // before scale / activation / ...; mean and rstd come from per-row partial sums left by the PRODUCER of the rows:
  const float* ln_stats = nullptr;    // [M][ln_chunks][2]: (sum, sum of squares) per 32-column chunk of the row
  const float* ln_colsum = nullptr;   // [N] sum over k of the (rounded) pre-multiplied weights
  int ln_chunks = 0;                  // chunks per row = C / 32
  float ln_eps = 0.f;
  // producer side: next to out (fp32) and out_bf16 (+ out_lo), emit the partial statistics of the OUTPUT rows
  float* stats_out = nullptr;         // [M][N / 32][2]
};

__device__ __forceinline__ float apply_epilogue(const Epilogue& e, float v, int row, int col, int N) {
  if (e.bias) v += __ldg(e.bias + col);
  if (e.scale) v = v * __ldg(e.scale + col) + __ldg(e.shift + col);
  if (e.act == ACT_RELU) v = fmaxf(v, 0.f);
  else if (e.act == ACT_GELU || e.act == ACT_GELU_EXACT) v = gelu_tanh(v);
  else if (e.act == ACT_SIGMOID) v = sigmoidf_(v);
  if (e.pos) v += __ldg(e.pos + (size_t)(row % e.pos_rows) * N + col);
  if (e.gate) v *= __ldg(e.gate + col);
  if (e.resid) {
    int rr = e.resid_mod ? (row % e.resid_mod) : row;
    v += __ldg(e.resid + (size_t)rr * e.ldr + col);
  }
  return v;
}

__device__ __forceinline__ void epilogue_dest(const Epilogue& e, int row, int col, int& orow, int& ocol) {
  if (e.group_rows) { orow = row % e.group_rows; ocol = col + (row / e.group_rows) * e.group_col_stride; }
  else { orow = row; ocol = col; }
}

enum { A_DIRECT = 0, A_CONV3 = 1 };
enum { B_NK = 0, B_KN = 1 };

struct GemmParams {
  const float* A = nullptr; long long lda = 0;
  int convH = 0, convW = 0, convC = 0;          // A_CONV3: NHWC input, K = 9*convC
  const float* B = nullptr; long long ldb = 0;
  int M = 0, N = 0, K = 0;
  int heads = 1, bdiv = 1;                      // blockIdx.z = n*heads + h ; B batch index = n / bdiv
  long long sAn = 0, sAh = 0, sBn = 0, sBh = 0, sCn = 0, sCh = 0;
  float alpha = 1.f;
  Epilogue e;
};

int gemm_f32(const GemmParams& p, int amode, int bmode, int batch, cudaStream_t st);

// convenience: C = epi(A * W^T), W K-major
inline GemmParams linear_params(const float* A, int lda, const float* W, int M, int N, int K) {
  GemmParams p; p.A = A; p.lda = lda; p.B = W; p.ldb = K; p.M = M; p.N = N; p.K = K; return p;
}

}  // namespace dvd
