#include "gemm_tc.cuh"
namespace dvd {
int gemm_tc_bf16(const __nv_bfloat16*, int, const __nv_bfloat16*, int, int, int, int, const Epilogue&, cudaStream_t) {
  set_error("gemm_tc_bf16: not built yet"); return DVD_E_BADARG;
}
int attention_tc_bf16(const __nv_bfloat16*, int, const __nv_bfloat16*, int, const __nv_bfloat16*, int, __nv_bfloat16*, int, int, int, int,
                      int, float, int, cudaStream_t) {
  set_error("attention_tc_bf16: not built yet"); return DVD_E_BADARG;
}
}
