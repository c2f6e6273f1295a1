// K11 — fused map-upsample + base ramp + affine + bilinear gather (HBM-bandwidth bound).
//
// Replaces, in ONE pass over the output, the reference chain
//   EV:301  F.interpolate(map64 -> HxW, bilinear, align_corners=True)
//   EV:304  F.interpolate(coords_grid(512,512)/511 -> HxW)          (the "base" identity ramp)
//   EV:306  ((s + base) * 2 - 1) * 0.987
//   visualization_utils.py:75 -> WP:73  F.grid_sample(photo, grid, bilinear, align_corners=True, zeros)
//   visualization_utils.py:76-77        .astype(uint8)               (u8 variants)
// Algorithmic HBM traffic: read photo once + write output once = 24 B/px (fp32 NCHW, C=3) or
// 6 B/px (uint8 HWC).  The reference moves ~88 B/px (two materialised fields, five elementwise
// passes, grid read, gather, store).
//
// Mapping: one thread = 4 consecutive output pixels of one row -> 128-bit coalesced stores per
// channel plane; a CTA covers a 128 x 8 pixel tile so that the gathered source footprint (the map
// is a smooth near-identity field) stays inside a few KB of L1.  The 64x64 coarse map (32 KB) is
// read through the read-only path and stays L1/L2 resident.
#include "common.cuh"

namespace dvd {

struct UnwarpGeom {
  int H, W, mh, mw;
  float sy, sx;      // (mh-1)/(H-1), (mw-1)/(W-1)   map upsample scales (align_corners=True)
  float by, bx;      // 511/(H-1), 511/(W-1)          base-ramp upsample scales
  float affine;      // 0.987
};

__host__ static UnwarpGeom make_geom(int H, int W, int mh, int mw, float affine) {
  UnwarpGeom g;
  g.H = H; g.W = W; g.mh = mh; g.mw = mw;
  // torch area_pixel_compute_scale(align_corners=True): (in-1)/(out-1) in fp32, 0 when out == 1
  g.sy = H > 1 ? (float)(mh - 1) / (float)(H - 1) : 0.f;
  g.sx = W > 1 ? (float)(mw - 1) / (float)(W - 1) : 0.f;
  g.by = H > 1 ? 511.0f / (float)(H - 1) : 0.f;
  g.bx = W > 1 ? 511.0f / (float)(W - 1) : 0.f;
  g.affine = affine;
  return g;
}

// bilinear weights of torch upsample_bilinear2d(align_corners=True) along one axis
__device__ __forceinline__ void up_coeff(float scale, int dst, int in_size, int& i0, int& ip, float& l0, float& l1) {
  float r = scale * (float)dst;
  i0 = (int)r;
  ip = (i0 < in_size - 1) ? 1 : 0;
  l1 = r - (float)i0;
  l0 = 1.0f - l1;
}

// value of the 512-ramp (k/511) upsampled to `dst` (EV:304 + EV:330-335 coords_grid_tensor)
__device__ __forceinline__ float base_ramp(float scale, int dst) {
  int k0, kp; float l0, l1;
  up_coeff(scale, dst, 512, k0, kp, l0, l1);
  float a = (float)k0 / 511.0f, b = (float)(k0 + kp) / 511.0f;
  return l0 * a + l1 * b;
}

// normalised sampling coordinates (gx, gy) in [-1,1]-ish for output pixel (i, j)
__device__ __forceinline__ void sample_coords(const UnwarpGeom& g, const float* __restrict__ map, int i, int j,
                                              float& gx, float& gy) {
  int y0, yp, x0, xp; float ly0, ly1, lx0, lx1;
  up_coeff(g.sy, i, g.mh, y0, yp, ly0, ly1);
  up_coeff(g.sx, j, g.mw, x0, xp, lx0, lx1);
  const float* m0 = map + (size_t)y0 * g.mw + x0;            // channel 0 = x displacement
  const float* m1 = m0 + (size_t)g.mh * g.mw;                // channel 1 = y displacement
  int dy = yp * g.mw;
  float sx = ly0 * (lx0 * __ldg(m0) + lx1 * __ldg(m0 + xp)) + ly1 * (lx0 * __ldg(m0 + dy) + lx1 * __ldg(m0 + dy + xp));
  float sy = ly0 * (lx0 * __ldg(m1) + lx1 * __ldg(m1 + xp)) + ly1 * (lx0 * __ldg(m1 + dy) + lx1 * __ldg(m1 + dy + xp));
  gx = ((sx + base_ramp(g.bx, j)) * 2.0f - 1.0f) * g.affine;
  gy = ((sy + base_ramp(g.by, i)) * 2.0f - 1.0f) * g.affine;
}

struct Taps {           // torch grid_sampler_2d bilinear, align_corners=True, padding zeros
  int x0, y0;
  float nw, ne, sw, se;
  bool vx0, vx1, vy0, vy1;
};

__device__ __forceinline__ Taps make_taps(float gx, float gy, int H, int W) {
  Taps t;
  float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
  float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
  float fx = floorf(ix), fy = floorf(iy);
  // clamp before the int conversion so that wild coordinates cannot overflow
  fx = fminf(fmaxf(fx, -2.f), (float)W + 1.f);
  fy = fminf(fmaxf(fy, -2.f), (float)H + 1.f);
  t.x0 = (int)fx; t.y0 = (int)fy;
  float ax = ix - fx, ay = iy - fy;       // == ix - ix_nw
  float bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;
  t.nw = bx * by; t.ne = ax * by; t.sw = bx * ay; t.se = ax * ay;
  t.vx0 = (t.x0 >= 0) & (t.x0 < W); t.vx1 = (t.x0 + 1 >= 0) & (t.x0 + 1 < W);
  t.vy0 = (t.y0 >= 0) & (t.y0 < H); t.vy1 = (t.y0 + 1 >= 0) & (t.y0 + 1 < H);
  return t;
}

template <typename T>
__device__ __forceinline__ float gather(const T* __restrict__ plane, const Taps& t, int W, int cstride) {
  // plane points at channel c of pixel (0,0); element (y,x) is plane[(y*W + x) * cstride]
  float v = 0.f;
  const T* p = plane + ((long long)t.y0 * W + t.x0) * cstride;
  if (t.vy0 & t.vx0) v += (float)__ldg(p) * t.nw;
  if (t.vy0 & t.vx1) v += (float)__ldg(p + cstride) * t.ne;
  if (t.vy1 & t.vx0) v += (float)__ldg(p + (long long)W * cstride) * t.sw;
  if (t.vy1 & t.vx1) v += (float)__ldg(p + (long long)(W + 1) * cstride) * t.se;
  return v;
}

__device__ __forceinline__ uint8_t to_u8_trunc(float v) {   // numpy .astype(uint8) on a non-negative value
  return (uint8_t)__float2uint_rz(fminf(fmaxf(v, 0.f), 255.f));
}

constexpr int TILE_W = 128, TILE_H = 8;   // 32 x 8 threads, 4 px per thread

// IN: 0 = fp32 NCHW, 1 = uint8 HWC.  OUT likewise.
template <int IN_U8, int OUT_U8, int C>
__global__ void __launch_bounds__(256) k_unwarp(const void* __restrict__ photo_, const float* __restrict__ map,
                                                void* __restrict__ out_, UnwarpGeom g) {
  const int b = blockIdx.z;
  const int i = blockIdx.y * TILE_H + threadIdx.y;
  const int j0 = (blockIdx.x * (TILE_W / 4) + threadIdx.x) * 4;
  if (i >= g.H || j0 >= g.W) return;
  const size_t plane = (size_t)g.H * g.W;
  const float* mapb = map + (size_t)b * 2 * g.mh * g.mw;
  float res[C][4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    int j = j0 + p;
    if (j < g.W) {
      float gx, gy;
      sample_coords(g, mapb, i, j, gx, gy);
      Taps t = make_taps(gx, gy, g.H, g.W);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (IN_U8) res[c][p] = gather((const uint8_t*)photo_ + (size_t)b * plane * C + c, t, g.W, C);
        else       res[c][p] = gather((const float*)photo_ + ((size_t)b * C + c) * plane, t, g.W, 1);
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) res[c][p] = 0.f;
    }
  }
  const bool full = (j0 + 3 < g.W);
  if (OUT_U8) {
    uint8_t* o = (uint8_t*)out_ + ((size_t)b * plane + (size_t)i * g.W + j0) * C;
    uint8_t bytes[4 * C];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < C; ++c) bytes[p * C + c] = to_u8_trunc(res[c][p]);
    if (full && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
#pragma unroll
      for (int w = 0; w < C; ++w) {   // 4*C bytes = C 32-bit words
        uint32_t v = bytes[4 * w] | (bytes[4 * w + 1] << 8) | (bytes[4 * w + 2] << 16) | ((uint32_t)bytes[4 * w + 3] << 24);
        reinterpret_cast<uint32_t*>(o)[w] = v;
      }
    } else {
      int n = min(4, g.W - j0) * C;
      for (int q = 0; q < n; ++q) o[q] = bytes[q];
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float* o = (float*)out_ + ((size_t)b * C + c) * plane + (size_t)i * g.W + j0;
      if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
        __stcs(reinterpret_cast<float4*>(o), make_float4(res[c][0], res[c][1], res[c][2], res[c][3]));  // streaming store
      } else {
        for (int p = 0; p < 4 && j0 + p < g.W; ++p) o[p] = res[c][p];
      }
    }
  }
}

__global__ void k_fullres_grid(const float* __restrict__ map, float* __restrict__ grid, UnwarpGeom g) {
  int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
  if (j >= g.W) return;
  float gx, gy;
  sample_coords(g, map + (size_t)b * 2 * g.mh * g.mw, i, j, gx, gy);
  size_t plane = (size_t)g.H * g.W;
  grid[((size_t)b * 2 + 0) * plane + (size_t)i * g.W + j] = gx;
  grid[((size_t)b * 2 + 1) * plane + (size_t)i * g.W + j] = gy;
}

// generic grid_sample (WP:73): img [B,C,H,W], grid [B,2,Ho,Wo]
__global__ void k_grid_sample(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
                              int C, int H, int W, int Ho, int Wo) {
  int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
  if (j >= Wo) return;
  size_t po = (size_t)Ho * Wo, pi = (size_t)H * W;
  float gx = __ldg(grid + ((size_t)b * 2 + 0) * po + (size_t)i * Wo + j);
  float gy = __ldg(grid + ((size_t)b * 2 + 1) * po + (size_t)i * Wo + j);
  Taps t = make_taps(gx, gy, H, W);
  for (int c = 0; c < C; ++c)
    out[((size_t)b * C + c) * po + (size_t)i * Wo + j] = gather(img + ((size_t)b * C + c) * pi, t, W, 1);
}

template <int IN_U8, int OUT_U8>
static int launch_unwarp(const void* photo, const float* map, void* out, int B, int C, int H, int W, int mh, int mw,
                         float affine, cudaStream_t st) {
  DVD_REQUIRE(B >= 0 && H >= 0 && W >= 0 && mh >= 1 && mw >= 1 && C >= 1 && C <= 4, "unwarp: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  if (B == 0 || H == 0 || W == 0) return 0;          // empty batch / empty photo: nothing to do
  DVD_REQUIRE(photo && map && out, "unwarp: null pointer");
  DVD_REQUIRE(B <= 65535 && cdiv(H, TILE_H) <= 65535, "unwarp: grid too large");
  UnwarpGeom g = make_geom(H, W, mh, mw, affine);
  dim3 grid(cdiv(W, TILE_W), cdiv(H, TILE_H), B), block(32, 8);
  switch (C) {
    case 1: k_unwarp<IN_U8, OUT_U8, 1><<<grid, block, 0, st>>>(photo, map, out, g); break;
    case 2: k_unwarp<IN_U8, OUT_U8, 2><<<grid, block, 0, st>>>(photo, map, out, g); break;
    case 3: k_unwarp<IN_U8, OUT_U8, 3><<<grid, block, 0, st>>>(photo, map, out, g); break;
    default: k_unwarp<IN_U8, OUT_U8, 4><<<grid, block, 0, st>>>(photo, map, out, g); break;
  }
  DVD_LAUNCH_CHECK("k_unwarp");
  return 0;
}

}  // namespace dvd

using namespace dvd;

extern "C" int dvd_unwarp_f32(const float* photo, const float* map, float* out, int B, int C, int H, int W, int mh, int mw,
                              float affine, void* stream) {
  return launch_unwarp<0, 0>(photo, map, out, B, C, H, W, mh, mw, affine, (cudaStream_t)stream);
}
extern "C" int dvd_unwarp_u8(const uint8_t* photo, const float* map, uint8_t* out, int B, int C, int H, int W, int mh, int mw,
                             float affine, void* stream) {
  return launch_unwarp<1, 1>(photo, map, out, B, C, H, W, mh, mw, affine, (cudaStream_t)stream);
}
extern "C" int dvd_unwarp_f32_u8(const float* photo, const float* map, uint8_t* out, int B, int C, int H, int W, int mh, int mw,
                                 float affine, void* stream) {
  return launch_unwarp<0, 1>(photo, map, out, B, C, H, W, mh, mw, affine, (cudaStream_t)stream);
}
extern "C" int dvd_fullres_grid_f32(const float* map, float* grid, int B, int H, int W, int mh, int mw, float affine,
                                    void* stream) {
  DVD_REQUIRE(map && grid, "fullres_grid: null pointer");
  if (B == 0 || H == 0 || W == 0) return 0;
  UnwarpGeom g = make_geom(H, W, mh, mw, affine);
  k_fullres_grid<<<dim3(cdiv(W, 256), H, B), 256, 0, (cudaStream_t)stream>>>(map, grid, g);
  DVD_LAUNCH_CHECK("k_fullres_grid");
  return 0;
}
extern "C" int dvd_grid_sample_f32(const float* img, const float* grid, float* out, int B, int C, int H, int W, int Ho, int Wo,
                                   void* stream) {
  DVD_REQUIRE(img && grid && out, "grid_sample: null pointer");
  if (B == 0 || C == 0 || Ho == 0 || Wo == 0) return 0;
  k_grid_sample<<<dim3(cdiv(Wo, 256), Ho, B), 256, 0, (cudaStream_t)stream>>>(img, grid, out, C, H, W, Ho, Wo);
  DVD_LAUNCH_CHECK("k_grid_sample");
  return 0;
}
