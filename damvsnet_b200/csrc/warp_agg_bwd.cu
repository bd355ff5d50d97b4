// Backward of the fused warp + aggregation kernel (training path, BatchNorm statistics held fixed).
//
// Mirrors what autograd does for reference models/cas_mvsnet.py:30-87: gradients flow to the reference
// feature (through the squared difference / the variance), to every source feature (through
// F.grid_sample: a bilinear scatter-add) and to the view-weight net's parameters; the sampling grid is
// not differentiated (models/module.py:307 builds it under no_grad) and neither are the hypotheses.
// Nothing of the forward is stored: each (pixel, depth, view) re-projects and re-samples, which costs
// less than reading back N x D warped volumes would.
//
// Mapping as in the forward kernel: a thread owns (pixel, 8 channels); the reference-feature gradient is
// accumulated in registers and written once, source-feature gradients are 16-byte vector atomics.
#include <cstdlib>

#include "warp_common.cuh"

namespace damvs {

struct WarpAggBwdParams {
  const float* ref;
  const float* src[kMaxSrcB];
  float* g_src[kMaxSrcB];
  const float* rot_trans;
  const float* hyp;
  const float* wnet;
  const void* g_vol;
  float* g_ref;
  float* g_wnet;
  int B, n_src, D, H, W, per_pixel;
};

template <int C, int MODE, typename GT>
__global__ void __launch_bounds__(128) warp_agg_bwd_kernel(const WarpAggBwdParams P) {
  constexpr int LPP = C / 8, PPW = 32 / LPP, TW = PPW, TH = 4;
  __shared__ float s_rt[kMaxSrcB * 12];
  __shared__ float s_wnet[C + 5];
  __shared__ float s_gw[C + 5];
  const int b = blockIdx.z, H = P.H, W = P.W, D = P.D, n_src = P.n_src;
  for (int i = threadIdx.x; i < n_src * 12; i += blockDim.x) s_rt[i] = P.rot_trans[((long long)(i / 12) * P.B + b) * 12 + i % 12];
  for (int i = threadIdx.x; i < C + 5; i += blockDim.x) {
    s_wnet[i] = (MODE == DAMVS_AGG_ADAPTIVE) ? P.wnet[i] : 0.f;
    s_gw[i] = 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane % LPP, pw = lane / LPP;
  const int px = blockIdx.x * TW + pw, py = blockIdx.y * TH + warp;
  const bool live = px < W && py < H;
  const int x = live ? px : 0, y = live ? py : 0, c0 = q * 8;
  const long long HW = (long long)H * W, img_stride = HW * C;
  const F8 rf = load8(P.ref + (long long)b * img_stride + ((long long)y * W + x) * C + c0);
  const float fx = (float)x, fy = (float)y, fw = (float)W, fh = (float)H;
  const float inv_half_w = 1.f / (float)((W - 1) / 2.0), inv_half_h = 1.f / (float)((H - 1) / 2.0);
  float w1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) w1[j] = s_wnet[c0 + j];
  const float s1 = s_wnet[C], b1 = s_wnet[C + 1], w2 = s_wnet[C + 2], s2 = s_wnet[C + 3], b2 = s_wnet[C + 4];
  const float* hyp = P.per_pixel ? P.hyp + (long long)b * D * HW + (long long)y * W + x : P.hyp + (long long)b * D;
  const long long hyp_stride = P.per_pixel ? HW : 1;
  const GT* gvol = reinterpret_cast<const GT*>(P.g_vol) + g8_offset(b, q, 0, y, x, C / 8, D, H, W);
  const long long vol_stride = HW * 8;
  const float inv_n = 1.f / (float)(n_src + 1), inv_nsrc = 1.f / (float)n_src;

  float gref[8], gw1[8];
  float gs1 = 0.f, gb1 = 0.f, gw2 = 0.f, gs2 = 0.f, gb2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) gref[j] = gw1[j] = 0.f;

  for (int d = 0; d < D; ++d) {
    const float dep = __ldg(hyp + d * hyp_stride);
    const F8 gv = load8(gvol + d * vol_stride);
    if (MODE == DAMVS_AGG_VARIANCE) {
      // vol = sum x^2 / n - (sum x / n)^2  =>  d vol / d x_i = 2 (x_i - mean) / n
      float mean[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) mean[j] = rf.v[j];
      for (int v = 0; v < n_src; ++v) {
        float ix, iy, w[4], wv[8];
        project_b(s_rt + v * 12, fx, fy, dep, inv_half_w, inv_half_h, fw, fh, ix, iy);
        const int off = footprint_b(ix, iy, H, W, C, w);
        blend8(P.src[v] + (long long)b * img_stride + off + c0, W, C, w, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) mean[j] += wv[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mean[j] *= inv_n;
        gref[j] += gv.v[j] * 2.f * (rf.v[j] - mean[j]) * inv_n;
      }
      for (int v = 0; v < n_src; ++v) {
        float ix, iy, w[4], wv[8], g[8];
        project_b(s_rt + v * 12, fx, fy, dep, inv_half_w, inv_half_h, fw, fh, ix, iy);
        const int off = footprint_b(ix, iy, H, W, C, w);
        blend8(P.src[v] + (long long)b * img_stride + off + c0, W, C, w, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = gv.v[j] * 2.f * (wv[j] - mean[j]) * inv_n;
        if (live) scatter8(P.g_src[v] + (long long)b * img_stride + off + c0, W, C, w, g);
      }
    } else {
      for (int v = 0; v < n_src; ++v) {
        float ix, iy, w[4], wv[8], e[8], df[8], g[8];
        project_b(s_rt + v * 12, fx, fy, dep, inv_half_w, inv_half_h, fw, fh, ix, iy);
        const int off = footprint_b(ix, iy, H, W, C, w);
        blend8(P.src[v] + (long long)b * img_stride + off + c0, W, C, w, wv);
        float s = 0.f, gwt = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          df[j] = rf.v[j] - wv[j];
          e[j] = df[j] * df[j];
          s = fmaf(w1[j], e[j], s);
          gwt = fmaf(gv.v[j] * inv_nsrc, e[j], gwt);   // d loss / d (wt + 1)
        }
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          gwt += __shfl_xor_sync(0xffffffffu, gwt, o);
        }
        const float pre1 = s * s1 + b1, a = fmaxf(pre1, 0.f);
        const float aw = a * w2, pre2 = aw * s2 + b2, wt1 = fmaxf(pre2, 0.f) + 1.f;
        const float gpre2 = pre2 > 0.f ? gwt : 0.f;
        const float ga = gpre2 * w2 * s2;
        const float gpre1 = pre1 > 0.f ? ga : 0.f;
        const float gs = gpre1 * s1;
        if (q == 0 && live) {  // scalars are identical on all lanes of the pixel: count them once
          gs2 += gpre2 * aw; gb2 += gpre2; gw2 += gpre2 * a * s2;
          gs1 += gpre1 * s; gb1 += gpre1;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float ge = gv.v[j] * inv_nsrc * wt1 + gs * w1[j];
          if (live) gw1[j] = fmaf(gs, e[j], gw1[j]);
          const float gdf = 2.f * df[j] * ge;
          gref[j] += gdf;
          g[j] = -gdf;
        }
        if (live) scatter8(P.g_src[v] + (long long)b * img_stride + off + c0, W, C, w, g);
      }
    }
  }
  if (live) {
    F8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) r.v[j] = gref[j];
    store8(P.g_ref + (long long)b * img_stride + ((long long)y * W + x) * C + c0, r);
  }
  if (MODE == DAMVS_AGG_ADAPTIVE && P.g_wnet) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_gw[c0 + j], gw1[j]);
    if (q == 0) {
      atomicAdd(&s_gw[C], gs1); atomicAdd(&s_gw[C + 1], gb1); atomicAdd(&s_gw[C + 2], gw2);
      atomicAdd(&s_gw[C + 3], gs2); atomicAdd(&s_gw[C + 4], gb2);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C + 5; i += blockDim.x)
      if (s_gw[i] != 0.f) atomicAdd(P.g_wnet + i, s_gw[i]);
  }
}

template <int C, int MODE>
static int launch_bwd(const WarpAggBwdParams& P, int g_dtype, cudaStream_t st) {
  constexpr int TW = 32 / (C / 8), TH = 4;
  dim3 grid((P.W + TW - 1) / TW, (P.H + TH - 1) / TH, P.B);
  if (g_dtype == DAMVS_F32) warp_agg_bwd_kernel<C, MODE, float><<<grid, 128, 0, st>>>(P);
  else warp_agg_bwd_kernel<C, MODE, __nv_bfloat16><<<grid, 128, 0, st>>>(P);
  DAMVS_LAUNCH_OK("warp_agg_bwd kernel");
  return DAMVS_OK;
}

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_warp_agg_bwd(const float* ref_nhwc, const float* const* src_nhwc, int n_src, const float* rot_trans,
                                  const float* depth_hyp, const float* wnet, const void* g_vol, int g_dtype, float* g_ref,
                                  float* const* g_src, float* g_wnet, int B, int C, int D, int H, int W, int mode,
                                  int per_pixel_hyp, void* stream) {
  DAMVS_REQUIRE(ref_nhwc && src_nhwc && rot_trans && depth_hyp && g_vol && g_ref && g_src, "warp_agg_bwd: null pointer");
  DAMVS_REQUIRE(n_src >= 1 && n_src <= kMaxSrcB, "warp_agg_bwd: n_src=%d outside [1,%d]", n_src, kMaxSrcB);
  DAMVS_REQUIRE(B > 0 && B <= 65535 && D > 0 && H > 1 && W > 1, "warp_agg_bwd: bad shape");
  DAMVS_REQUIRE(mode == DAMVS_AGG_VARIANCE || (mode == DAMVS_AGG_ADAPTIVE && wnet), "warp_agg_bwd: bad mode / missing wnet");
  DAMVS_REQUIRE(g_dtype == DAMVS_F32 || g_dtype == DAMVS_BF16, "warp_agg_bwd: bad g_dtype");
  WarpAggBwdParams P;
  P.ref = ref_nhwc;
  for (int v = 0; v < kMaxSrcB; ++v) { P.src[v] = v < n_src ? src_nhwc[v] : nullptr; P.g_src[v] = v < n_src ? g_src[v] : nullptr; }
  for (int v = 0; v < n_src; ++v)
    DAMVS_REQUIRE(P.src[v] && P.g_src[v] && aligned16(P.src[v]) && aligned16(P.g_src[v]), "warp_agg_bwd: src/g_src[%d] null or misaligned", v);
  DAMVS_REQUIRE(aligned16(ref_nhwc) && aligned16(g_ref) && aligned16(g_vol), "warp_agg_bwd: pointers must be 16-byte aligned");
  P.rot_trans = rot_trans; P.hyp = depth_hyp; P.wnet = wnet; P.g_vol = g_vol; P.g_ref = g_ref; P.g_wnet = g_wnet;
  P.B = B; P.n_src = n_src; P.D = D; P.H = H; P.W = W; P.per_pixel = per_pixel_hyp;
  cudaStream_t st = (cudaStream_t)stream;
#define DISPATCH(CC)                                                                        \
  case CC:                                                                                  \
    return mode == DAMVS_AGG_ADAPTIVE ? launch_bwd<CC, DAMVS_AGG_ADAPTIVE>(P, g_dtype, st) \
                                      : launch_bwd<CC, DAMVS_AGG_VARIANCE>(P, g_dtype, st);
  switch (C) {
    DISPATCH(8)
    DISPATCH(16)
    DISPATCH(32)
    DISPATCH(64)
    default:
      return set_error(DAMVS_ERR_UNSUPPORTED, "warp_agg_bwd: C=%d not in {8,16,32,64}", C);
  }
#undef DISPATCH
}
