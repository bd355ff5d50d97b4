// Depth-hypothesis sampling for cascade stages 2 and 3, fused.
//
// Replaces, for `depth is not None` (reference models/cas_mvsnet.py:250-253, 269-274, 293-296 and
// uncertainty_aware_samples, models/module.py:999-1038):
//     cur  = bilinear_upsample(prev_depth, full res, align_corners=False)
//     ev   = bilinear_upsample(prev_variance, full res)
//     low  = -min(cur, ev);  step = (ev - low) / (D - 1)
//     w    = softmax_i( 3 (low + step i) / (ev + eps) )
//     s_i  = cur + low + step i + eps + w_i step                      (full resolution, D planes)
//     out  = trilinear_resample(s, [D, H/scale, W/scale], align_corners=False)
// The reference materialises D full-resolution planes (as Python lists of D tensors) plus ~5 temporaries of
// that size and then down-samples; here one thread owns one OUTPUT pixel, rebuilds the 1 (scale 1) or 4 (scale
// 2, 4: with align_corners=False an exact 1/scale resample is the mean of the centre 2x2 block) full-resolution
// hypothesis columns in registers and writes D values.  HBM traffic: two small maps read, D*h*w*4 bytes written.
// The soft-max needs no max pass: the logits increase with i (step >= 0), so the maximum is the last one.
#include "common.cuh"

namespace damvs {

constexpr float kEpsHyp = 1e-12f;  // reference models/module.py:10

// F.interpolate(mode='bilinear', align_corners=False): source index of destination index `d`
__device__ __forceinline__ void lin_src(int d, float scale, int n, int& i0, int& i1, float& lam) {
  float s = fmaxf(((float)d + 0.5f) * scale - 0.5f, 0.f);
  i0 = min((int)s, n - 1);
  i1 = min(i0 + 1, n - 1);
  lam = s - (float)i0;
}

__device__ __forceinline__ float bilerp(const float* __restrict__ m, int wp, int y0, int y1, float ly, int x0, int x1, float lx) {
  const float a = __ldg(m + (long long)y0 * wp + x0), b = __ldg(m + (long long)y0 * wp + x1);
  const float c = __ldg(m + (long long)y1 * wp + x0), d = __ldg(m + (long long)y1 * wp + x1);
  // ATen upsample_bilinear2d: h0lambda * (w0lambda * a + w1lambda * b) + h1lambda * (w0lambda * c + w1lambda * d)
  return (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * c + lx * d);
}

__global__ void __launch_bounds__(256) uncertainty_samples_kernel(const float* __restrict__ prev_depth, const float* __restrict__ prev_var,
                                                                  float* __restrict__ out, int hp, int wp, int D, int H, int W, int h, int w,
                                                                  int scale) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const float* pd = prev_depth + (long long)b * hp * wp;
  const float* pv = prev_var + (long long)b * hp * wp;
  const float sy = (float)hp / (float)H, sx = (float)wp / (float)W;
  const int nsub = scale == 1 ? 1 : 2;
  const int fy0 = scale == 1 ? y : y * scale + scale / 2 - 1, fx0 = scale == 1 ? x : x * scale + scale / 2 - 1;
  float cur[4], low[4], step[4], kk[4], mx[4], inv_sum[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (s >= nsub * nsub) { cur[s] = low[s] = step[s] = kk[s] = mx[s] = inv_sum[s] = 0.f; continue; }
    const int fy = min(fy0 + s / nsub, H - 1), fx = min(fx0 + s % nsub, W - 1);
    int y0, y1, x0, x1;
    float ly, lx;
    lin_src(fy, sy, hp, y0, y1, ly);
    lin_src(fx, sx, wp, x0, x1, lx);
    const float c = bilerp(pd, wp, y0, y1, ly, x0, x1, lx), ev = bilerp(pv, wp, y0, y1, ly, x0, x1, lx);
    const float lo = -fminf(c, ev);
    const float st = (ev - lo) / ((float)D - 1.f);
    cur[s] = c; low[s] = lo; step[s] = st;
    kk[s] = 3.f / (ev + kEpsHyp);
    mx[s] = fmaxf(lo * kk[s], (lo + st * (float)(D - 1)) * kk[s]);
    float sum = 0.f;
    for (int i = 0; i < D; ++i) sum += __expf((lo + st * (float)i) * kk[s] - mx[s]);
    inv_sum[s] = 1.f / sum;
  }
  const float wsub = 1.f / (float)(nsub * nsub);
  float* o = out + (long long)b * D * h * w + (long long)y * w + x;
  for (int i = 0; i < D; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s < nsub * nsub) {
        const float lin = low[s] + step[s] * (float)i;
        const float off = __expf(lin * kk[s] - mx[s]) * inv_sum[s];
        acc += (cur[s] + lin + kEpsHyp + off * step[s]) * wsub;
      }
    }
    o[(long long)i * h * w] = acc;
  }
}

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_uncertainty_samples_fwd(const float* prev_depth, const float* prev_var, float* out, int B, int hp, int wp,
                                             int D, int H, int W, int scale, void* stream) {
  DAMVS_REQUIRE(prev_depth && prev_var && out, "uncertainty_samples: null pointer");
  DAMVS_REQUIRE(B > 0 && B <= 65535 && hp > 0 && wp > 0 && D > 1 && H > 0 && W > 0, "uncertainty_samples: bad shape (ndepth must be > 1)");
  DAMVS_REQUIRE(scale == 1 || scale == 2 || scale == 4, "uncertainty_samples: scale %d not in {1,2,4}", scale);
  DAMVS_REQUIRE(H % scale == 0 && W % scale == 0, "uncertainty_samples: H, W must be multiples of the stage scale");
  const int h = H / scale, w = W / scale;
  DAMVS_REQUIRE(h <= 65535, "uncertainty_samples: too many rows");
  dim3 grid((w + 255) / 256, h, B);
  uncertainty_samples_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(prev_depth, prev_var, out, hp, wp, D, H, W, h, w, scale);
  DAMVS_LAUNCH_OK("uncertainty_samples kernel");
  return DAMVS_OK;
}
