// Shared host/device helpers for libdvd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dvd_b200.h"

namespace dvd {

// ---- error plumbing (thread-local message, integer codes across the C ABI) ----
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define DVD_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::dvd::cuda_fail(_e, #expr); \
  } while (0)

// call after every kernel launch
#define DVD_LAUNCH_CHECK(name)                                              \
  do {                                                                      \
    ::dvd::count_launch();                                                  \
    cudaError_t _e = cudaGetLastError();                                    \
    if (_e != cudaSuccess) return ::dvd::cuda_fail(_e, "launch " name);     \
  } while (0)

#define DVD_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) {                                \
      ::dvd::set_error(__VA_ARGS__);              \
      return DVD_E_BADARG;                        \
    }                                             \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kSMs = 148;            // B200: 2 dies x 74 SMs
constexpr int kTokens = 1024;        // 32 x 32 patches of the 64 x 64 map
constexpr int kHid = 384;            // DiT-S hidden
constexpr int kDec = 1536;           // decoder d_model
constexpr int kDecInner = 2048;
constexpr int kHeads = 6;

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  // torch GELU(approximate='tanh'): 0.5 x (1 + tanh( sqrt(2/pi) (x + 0.044715 x^3) ))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}
// hardware tanh (MUFU.TANH, ~2^-11 relative error): used by the bf16 tensor-path epilogue only
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x), t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace dvd
