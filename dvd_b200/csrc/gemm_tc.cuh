// tcgen05 / TMEM / TMA tensor-core path (DVD_PREC_BF16): dense GEMM with the shared Epilogue and
// flash attention.  Implemented in gemm_tc.cu / attn_tc.cu.
#pragma once
#include "gemm_simt.cuh"

namespace dvd {

// C[M,N] = epilogue(A[M,K] * W[N,K]^T), A/W bf16 row-major (K contiguous), fp32 accumulate in TMEM.
int gemm_tc_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K, const Epilogue& e,
                 cudaStream_t st);

// 3x3 / pad 1 conv as implicit GEMM on tcgen05 (TMA boxes over the NHWC activation, zero padding by OOB fill):
// in [B,H,W,Cin] bf16, Wt [Cout, 9*Cin] bf16 ordered [ky][kx][Cin]; output through the Epilogue as [B*H*W, Cout].
int conv3x3_tc_bf16(const __nv_bfloat16* in, const __nv_bfloat16* Wt, int B, int H, int Wd, int Cin, int Cout, const Epilogue& e,
                    cudaStream_t st);

// softmax(scale * Q K^T) V per (sample, head); bf16 in/out, fp32 softmax statistics and accumulation.
// q/k/o row-major with leading dims ld*, head h at columns [h*d, (h+1)*d); vt is V TRANSPOSED: [nsamp/kv_div, heads*d, T]
// (written by the QKV GEMM epilogue, Epilogue::vt_out).  k/vt of sample n come from sample n / kv_div.
// d in {64, 256}, T multiple of 128.
int attention_tc_bf16(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* vt, __nv_bfloat16* o, int ldo,
                      int nsamp, int heads, int T, int d, float scale, int kv_div, cudaStream_t st);

// Same, for up to 4 key/value contexts that share the queries (one launch): context i uses k[i], vt[i], kv_div[i], writes o[i].
int attention_tc_bf16_multi(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* const* k, int ldk, const __nv_bfloat16* const* vt,
                            __nv_bfloat16* const* o, int ldo, const int* kv_div, int nctx, int nsamp, int heads, int T, int d, float scale,
                            cudaStream_t st);

// V [nsamp, T, C] (row stride ldv) -> V^T [nsamp, C, T]  (test hook only)
int transpose_v_bf16(const __nv_bfloat16* v, int ldv, __nv_bfloat16* vt, int nsamp, int T, int C, cudaStream_t st);

}  // namespace dvd
