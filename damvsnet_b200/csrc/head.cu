// Softmax / depth-regression / photometric-confidence / variance head.
//
// Replaces reference models/cas_mvsnet.py:105-124 and depth_regression
// (models/module.py:609-615): one pass over the logits and hypotheses, one
// store of the probability volume and three [B,H,W] maps.  HBM-bound streaming:
// algorithmic bytes per pixel = 2*D*4 read + D*4 + 12 written.
//
// Mapping: one thread per pixel, the whole D column in registers (D <= 64 for
// the register path); lanes run along W so every load/store of plane k is a
// coalesced 128-byte line per warp and a thread has D independent loads in
// flight.  A column is strided by H*W in memory, so there is nothing to stage
// in shared memory and nothing to reduce across lanes.
#include "common.cuh"

namespace damvs {

template <int MAXD>
__global__ void __launch_bounds__(256) head_reg_kernel(const float* __restrict__ logits,
                                                       const float* __restrict__ hyp, float* __restrict__ prob,
                                                       float* __restrict__ depth, float* __restrict__ conf,
                                                       float* __restrict__ var, int D, long long HW,
                                                       long long total, int per_pixel) {
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  long long b = pix / HW;
  long long p = pix - b * HW;
  const float* lg = logits + b * D * HW + p;
  float e[MAXD];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < MAXD; ++k) {
    if (k < D) {
      e[k] = __ldcs(lg + (long long)k * HW);
      m = fmaxf(m, e[k]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXD; ++k) {
    if (k < D) {
      e[k] = expf(e[k] - m);
      s += e[k];
    }
  }
  const float* hp = per_pixel ? hyp + b * D * HW + p : hyp + b * D;
  const long long hs = per_pixel ? HW : 1;
  float dsum = 0.f, isum = 0.f, d2sum = 0.f;
  float* pr = prob ? prob + b * D * HW + p : nullptr;
  // first sweep: probabilities, expected depth and expected index
#pragma unroll
  for (int k = 0; k < MAXD; ++k) {
    if (k < D) {
      float pk = e[k] / s;
      e[k] = pk;
      float dk = __ldg(hp + k * hs);
      dsum += pk * dk;
      isum += pk * (float)k;
      if (pr) __stcs(pr + (long long)k * HW, pk);
    }
  }
  // second sweep: hypothesis variance about the expected depth (hypotheses are L1/L2 hits)
#pragma unroll
  for (int k = 0; k < MAXD; ++k) {
    if (k < D) {
      float dk = __ldg(hp + k * hs);
      float t = dk - dsum;
      d2sum += (t * t) * e[k];
    }
  }
  long long idx = (long long)isum;  // .long() truncation, reference cas_mvsnet.py:116
  idx = idx < 0 ? 0 : (idx > D - 1 ? D - 1 : idx);
  float c = 0.f;
#pragma unroll
  for (int k = 0; k < MAXD; ++k) {
    if (k < D) {
      if (k >= idx - 1 && k <= idx + 2) c += e[k];
    }
  }
  depth[pix] = dsum;
  conf[pix] = c;
  var[pix] = 3.f * sqrtf(d2sum);
}

// Any D: three passes over the logits column (re-reads are L2 hits).
__global__ void __launch_bounds__(256) head_stream_kernel(const float* __restrict__ logits,
                                                          const float* __restrict__ hyp, float* __restrict__ prob,
                                                          float* __restrict__ depth, float* __restrict__ conf,
                                                          float* __restrict__ var, int D, long long HW,
                                                          long long total, int per_pixel) {
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  long long b = pix / HW;
  long long p = pix - b * HW;
  const float* lg = logits + b * D * HW + p;
  const float* hp = per_pixel ? hyp + b * D * HW + p : hyp + b * D;
  const long long hs = per_pixel ? HW : 1;
  float m = -INFINITY;
  for (int k = 0; k < D; ++k) m = fmaxf(m, __ldg(lg + (long long)k * HW));
  float s = 0.f;
  for (int k = 0; k < D; ++k) s += expf(__ldg(lg + (long long)k * HW) - m);
  float dsum = 0.f, isum = 0.f;
  float* pr = prob ? prob + b * D * HW + p : nullptr;
  for (int k = 0; k < D; ++k) {
    float pk = expf(__ldg(lg + (long long)k * HW) - m) / s;
    dsum += pk * __ldg(hp + k * hs);
    isum += pk * (float)k;
    if (pr) pr[(long long)k * HW] = pk;
  }
  long long idx = (long long)isum;
  idx = idx < 0 ? 0 : (idx > D - 1 ? D - 1 : idx);
  float d2sum = 0.f, c = 0.f;
  for (int k = 0; k < D; ++k) {
    float pk = expf(__ldg(lg + (long long)k * HW) - m) / s;
    float t = __ldg(hp + k * hs) - dsum;
    d2sum += (t * t) * pk;
    if (k >= idx - 1 && k <= idx + 2) c += pk;
  }
  depth[pix] = dsum;
  conf[pix] = c;
  var[pix] = 3.f * sqrtf(d2sum);
}

__global__ void __launch_bounds__(256) depth_regression_kernel(const float* __restrict__ prob,
                                                               const float* __restrict__ hyp,
                                                               float* __restrict__ out, int D, long long HW,
                                                               long long total, int per_pixel) {
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  long long b = pix / HW;
  long long p = pix - b * HW;
  const float* pp = prob + b * D * HW + p;
  const float* hp = per_pixel ? hyp + b * D * HW + p : hyp + b * D;
  const long long hs = per_pixel ? HW : 1;
  float acc = 0.f;
  for (int k = 0; k < D; ++k) acc += __ldg(pp + (long long)k * HW) * __ldg(hp + k * hs);
  out[pix] = acc;
}

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_softmax_regress_fwd(const float* logits, const float* depth_hyp, float* prob, float* depth,
                                         float* conf, float* var, int B, int D, int H, int W, int per_pixel_hyp,
                                         void* stream) {
  DAMVS_REQUIRE(logits && depth_hyp && depth && conf && var, "softmax_regress: null pointer");
  DAMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "softmax_regress: bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
  long long HW = (long long)H * W, total = HW * B;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned blocks = (unsigned)((total + 255) / 256);
#define LAUNCH(MAXD) \
  head_reg_kernel<MAXD><<<blocks, 256, 0, st>>>(logits, depth_hyp, prob, depth, conf, var, D, HW, total, per_pixel_hyp)
  if (D <= 8) LAUNCH(8);
  else if (D <= 16) LAUNCH(16);
  else if (D <= 32) LAUNCH(32);
  else if (D <= 48) LAUNCH(48);
  else if (D <= 64) LAUNCH(64);
  else
    head_stream_kernel<<<blocks, 256, 0, st>>>(logits, depth_hyp, prob, depth, conf, var, D, HW, total,
                                               per_pixel_hyp);
#undef LAUNCH
  DAMVS_LAUNCH_OK("head kernel");
  return DAMVS_OK;
}

extern "C" int damvs_depth_regression_fwd(const float* prob, const float* depth_hyp, float* out, int B, int D,
                                          int H, int W, int per_pixel_hyp, void* stream) {
  DAMVS_REQUIRE(prob && depth_hyp && out, "depth_regression: null pointer");
  DAMVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "depth_regression: bad shape");
  long long HW = (long long)H * W, total = HW * B;
  unsigned blocks = (unsigned)((total + 255) / 256);
  depth_regression_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(prob, depth_hyp, out, D, HW, total,
                                                                     per_pixel_hyp);
  DAMVS_LAUNCH_OK("depth_regression kernel");
  return DAMVS_OK;
}
