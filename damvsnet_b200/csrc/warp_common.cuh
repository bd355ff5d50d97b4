// Device helpers shared by the training-path warp kernels (warp_agg_bwd.cu, warp_agg_train.cu): projection of a
// reference pixel at one hypothesis (reference models/module.py:312-325, operation order kept), the "always
// addressable" bilinear footprint (zero padding per tap), 8-channel blend and 8-channel vector-atomic scatter.
#pragma once
#include "common.cuh"

namespace damvs {

constexpr int kMaxSrcB = 15;

__device__ __forceinline__ void project_b(const float* rt, float fx, float fy, float d, float inv_half_w, float inv_half_h, float fw,
                                          float fh, float& ix, float& iy) {
  const float rx = fmaf(rt[0], fx, fmaf(rt[1], fy, rt[2]));
  const float ry = fmaf(rt[3], fx, fmaf(rt[4], fy, rt[5]));
  const float rz = fmaf(rt[6], fx, fmaf(rt[7], fy, rt[8]));
  const float px = __fadd_rn(__fmul_rn(rx, d), rt[9]);
  const float py = __fadd_rn(__fmul_rn(ry, d), rt[10]);
  const float pz = __fadd_rn(__fmul_rn(rz, d), rt[11]);
  const float u = __fdiv_rn(px, pz), v = __fdiv_rn(py, pz);
  const float gx = __fadd_rn(__fmul_rn(u, inv_half_w), -1.f), gy = __fadd_rn(__fmul_rn(v, inv_half_h), -1.f);
  ix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), fw), -1.f), 0.5f);
  iy = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), fh), -1.f), 0.5f);
}

// clamped 2x2 block + weights with out-of-range taps zeroed (same construction as the forward kernel)
__device__ __forceinline__ int footprint_b(float ix, float iy, int H, int W, int C, float (&w)[4]) {
  w[0] = w[1] = w[2] = w[3] = 0.f;
  if (!(ix > -1.f && ix < (float)W && iy > -1.f && iy < (float)H)) return 0;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  float wl = (fx0 + 1.f) - ix, wr = ix - fx0, wt = (fy0 + 1.f) - iy, wb = iy - fy0;
  int xc = x0, yc = y0;
  if (x0 < 0) { xc = 0; wl = wr; wr = 0.f; } else if (x0 > W - 2) { xc = W - 2; wr = wl; wl = 0.f; }
  if (y0 < 0) { yc = 0; wt = wb; wb = 0.f; } else if (y0 > H - 2) { yc = H - 2; wb = wt; wt = 0.f; }
  w[0] = wl * wt; w[1] = wr * wt; w[2] = wl * wb; w[3] = wr * wb;
  return (yc * W + xc) * C;
}

__device__ __forceinline__ void blend8(const float* p, int W, int C, const float (&w)[4], float (&o)[8]) {
  const F8 a = load8(p), b = load8(p + C), c = load8(p + (long long)W * C), d = load8(p + (long long)W * C + C);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaf(d.v[j], w[3], fmaf(c.v[j], w[2], fmaf(b.v[j], w[1], a.v[j] * w[0])));
}

__device__ __forceinline__ void scatter8(float* p, int W, int C, const float (&w)[4], const float (&g)[8]) {
  float* q[4] = {p, p + C, p + (long long)W * C, p + (long long)W * C + C};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (w[k] != 0.f) {
      atomicAdd(reinterpret_cast<float4*>(q[k]), make_float4(g[0] * w[k], g[1] * w[k], g[2] * w[k], g[3] * w[k]));
      atomicAdd(reinterpret_cast<float4*>(q[k]) + 1, make_float4(g[4] * w[k], g[5] * w[k], g[6] * w[k], g[7] * w[k]));
    }
  }
}

}  // namespace damvs
