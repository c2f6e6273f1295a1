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

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device): function attributes are per device
// context, so a per-process flag would leave the second GPU of a process without the opt-in.
#define DVD_SET_MAX_SMEM(kernel, bytes)                                                                   \
  do {                                                                                                    \
    static bool _set[64] = {};                                                                            \
    int _dev = 0;                                                                                         \
    DVD_CUDA(cudaGetDevice(&_dev));                                                                       \
    if (_dev < 0 || _dev >= 64 || !_set[_dev]) {                                                          \
      DVD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));  \
      if (_dev >= 0 && _dev < 64) _set[_dev] = true;                                                      \
    }                                                                                                     \
  } while (0)

constexpr int kTokens = 1024;        // 32 x 32 patches of the 64 x 64 map
constexpr int kHid = 384;            // DiT-S hidden
constexpr int kDec = 1536;           // decoder d_model
constexpr int kDecInner = 2048;
constexpr int kHeads = 6;

// ---- programmatic dependent launch (PDL) ----
// Kernels launched through launch_pdl() may begin (prologue: barrier init, TMEM allocation, descriptor prefetch) while the
// previous kernel of the stream is still draining; they call pdl_wait() before touching any global memory the predecessor
// may have written, and pdl_trigger() at their start so that THEIR successor can be scheduled early.  Set DVD_NO_PDL=1 to
// fall back to plain stream serialisation.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled(int kind = 0x1F);    // kind: 1 = GEMM, 2 = attention, 4 = layer norm, 8 = depthwise conv, 16 = small dense layers (DVD_PDL_MASK)
bool debug_skip(int kind);            // DVD_DEBUG_SKIP=<mask of kinds>: timing ablation, see api.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  if (debug_skip(kind)) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(kind) ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Same, for a kernel launched as thread-block clusters of (clx, cly, 1) CTAs.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int clx, int cly,
                                      Args... args) {
  if (debug_skip(kind)) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(kind) ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = clx; attr[1].val.clusterDim.y = cly; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
