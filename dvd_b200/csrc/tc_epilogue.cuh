// Output side of the tensor-core epilogues (shared by gemm_tc.cu, gemm_pair.cu and attn_tc.cu): fp32 and / or 16-bit stores of
// four consecutive columns.  16-bit flavours: bf16 (DVD_PREC_BF16), bf16 hi + lo pair (operand of a 3-pass split-precision GEMM,
// DVD_PREC_BF16X3) or fp16 (operand of the fp16 attention kernel in DVD_PREC_BF16X3).
#pragma once
#include <cuda_fp16.h>
#include "gemm_simt.cuh"

namespace dvd {

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  __half2 p = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// hi = bf16(v), lo = bf16(v - hi): v == hi + lo to 2^-17 relative
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = pack_bf16x2(a - __low2float(h), b - __high2float(h));
}
__device__ __forceinline__ uint16_t cvt16(float v, int f16) {
  if (f16) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// v[0..3] -> e.out / e.out_bf16 (+ e.out_lo) at (orow, ocol .. ocol+3)
__device__ __forceinline__ void store_tc_out4(const Epilogue& e, int orow, int ocol, const float (&v)[4]) {
  if (e.out) *reinterpret_cast<float4*>(e.out + (size_t)orow * e.ldc + ocol) = make_float4(v[0], v[1], v[2], v[3]);
  if (e.out_bf16) {
    const size_t off = (size_t)orow * e.ldc_bf16 + ocol;
    uint2 u;
    if (e.out_lo) {
      uint2 l;
      split_bf16x2(v[0], v[1], u.x, l.x);
      split_bf16x2(v[2], v[3], u.y, l.y);
      *reinterpret_cast<uint2*>(e.out_lo + off) = l;
    } else if (e.out_f16) {
      u.x = pack_f16x2(v[0], v[1]); u.y = pack_f16x2(v[2], v[3]);
    } else {
      u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    }
    *reinterpret_cast<uint2*>(e.out_bf16 + off) = u;
  }
}

// the fused epilogue on four consecutive columns: bias, folded BN, activation, pos-embed, gate, residual (in this order, like
// apply_epilogue of the fp32 path).  cb/cs/ct/cg are the per-column vectors of this lane, q / p the residual / pos-embed values.
struct EpiCols {
  float4 cb, cs, ct, cg;
  bool has_scale, has_gate;
  int act;
};
__device__ __forceinline__ EpiCols load_epi_cols(const Epilogue& e, int col) {
  EpiCols c;
  c.cb = make_float4(0.f, 0.f, 0.f, 0.f); c.cs = make_float4(1.f, 1.f, 1.f, 1.f); c.ct = c.cb; c.cg = c.cs;
  if (e.bias) c.cb = __ldg(reinterpret_cast<const float4*>(e.bias + col));
  if (e.scale) { c.cs = __ldg(reinterpret_cast<const float4*>(e.scale + col)); c.ct = __ldg(reinterpret_cast<const float4*>(e.shift + col)); }
  if (e.gate) c.cg = __ldg(reinterpret_cast<const float4*>(e.gate + col));
  c.has_scale = e.scale != nullptr; c.has_gate = e.gate != nullptr; c.act = e.act;
  return c;
}
__device__ __forceinline__ void apply_epi4(const EpiCols& c, const float4& a, bool has_pos, const float4& p, bool has_res, const float4& q,
                                           float (&v)[4]) {
  v[0] = a.x + c.cb.x; v[1] = a.y + c.cb.y; v[2] = a.z + c.cb.z; v[3] = a.w + c.cb.w;
  if (c.has_scale) { v[0] = v[0] * c.cs.x + c.ct.x; v[1] = v[1] * c.cs.y + c.ct.y; v[2] = v[2] * c.cs.z + c.ct.z; v[3] = v[3] * c.cs.w + c.ct.w; }
  if (c.act == ACT_RELU) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
  else if (c.act == ACT_GELU) { v[0] = gelu_tanh_fast(v[0]); v[1] = gelu_tanh_fast(v[1]); v[2] = gelu_tanh_fast(v[2]); v[3] = gelu_tanh_fast(v[3]); }
  else if (c.act == ACT_GELU_EXACT) { v[0] = gelu_tanh(v[0]); v[1] = gelu_tanh(v[1]); v[2] = gelu_tanh(v[2]); v[3] = gelu_tanh(v[3]); }
  else if (c.act == ACT_SIGMOID) { v[0] = sigmoidf_(v[0]); v[1] = sigmoidf_(v[1]); v[2] = sigmoidf_(v[2]); v[3] = sigmoidf_(v[3]); }
  if (has_pos) { v[0] += p.x; v[1] += p.y; v[2] += p.z; v[3] += p.w; }
  if (c.has_gate) { v[0] *= c.cg.x; v[1] *= c.cg.y; v[2] *= c.cg.z; v[3] *= c.cg.w; }
  if (has_res) { v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
}

}  // namespace dvd
