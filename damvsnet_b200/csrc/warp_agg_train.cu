// Warp + aggregation with the view-weight net in TRAINING mode (batch statistics).
//
// In eval mode AggWeightNetVolume (reference models/module.py:544-563) is a per-voxel function and is fused
// into warp_agg_kernel.  In training mode its two BatchNorm3d layers normalise over every voxel of a source
// view's [B,1,D,h,w] score volume (models/cas_mvsnet.py:71 calls the net once per source view, so its
// statistics -- and its running buffers -- are per view), which makes the weight non-local.  The adaptive
// aggregation then factors into
//     s_v   = sum_c w1[c] (ref - warp_v)[c]^2                    "score"    (this file, native)
//     wt_v  = relu(bn2(w2 * relu(bn1(s_v))))                     scalar volumes, 1/C of the data (wnet_chain.cu)
//     vol   = sum_v (wt_v + 1) (ref - warp_v)^2 / n_src          "weighted" (this file)
// and the backward into: d loss / d wt_v without any scatter ("gwt"), the chain's backward, and ONE scatter pass that
// carries both paths into the features ("merged").  Every kernel re-projects and re-samples instead of storing the
// N x D warped volumes.  Thread = (pixel, 8 channels), the C/8 lanes of a pixel adjacent in the warp.
#include "warp_common.cuh"

namespace damvs {

enum { OP_SCORE_FWD = 0, OP_WEIGHTED_FWD = 1, OP_GWT = 2, OP_MERGED_BWD = 3 };

struct WarpTrainParams {
  const float* ref;
  const float* src[kMaxSrcB];
  float* g_src[kMaxSrcB];
  const float* rot_trans;
  const float* hyp;
  const float* w1;       // [C]                       (score ops)
  const float* wt_vol;   // [n_src][B][D][H][W]       (weighted ops)
  float* s_vol;          // score fwd: out; gwt: g_wt out; merged bwd: g_s in
  void* vol;             // weighted fwd: out (G8); gwt / merged bwd: g_vol in (G8)
  float* g_ref;          // accumulated
  float* g_w1;           // accumulated [C]
  int B, n_src, D, H, W, per_pixel;
};

template <int C, int OP, typename VT>
__global__ void __launch_bounds__(128) warp_train_kernel(const WarpTrainParams P) {
  constexpr int LPP = C / 8, PPW = 32 / LPP, TW = PPW, TH = 4;
  __shared__ float s_rt[kMaxSrcB * 12];
  __shared__ float s_w1[C];
  __shared__ float s_gw[C];
  const int b = blockIdx.z, H = P.H, W = P.W, D = P.D, n_src = P.n_src;
  for (int i = threadIdx.x; i < n_src * 12; i += blockDim.x) s_rt[i] = P.rot_trans[((long long)(i / 12) * P.B + b) * 12 + i % 12];
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_w1[i] = (OP == OP_SCORE_FWD || OP == OP_MERGED_BWD) ? P.w1[i] : 0.f;
    s_gw[i] = 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane % LPP, pw = lane / LPP;
  const int px = blockIdx.x * TW + pw, py = blockIdx.y * TH + warp;
  const bool live = px < W && py < H;
  const int x = live ? px : 0, y = live ? py : 0, c0 = q * 8;
  const long long HW = (long long)H * W, img_stride = HW * C, DHW = HW * D;
  const F8 rf = load8(P.ref + (long long)b * img_stride + ((long long)y * W + x) * C + c0);
  const float fx = (float)x, fy = (float)y, fw = (float)W, fh = (float)H;
  const float inv_half_w = 1.f / (float)((W - 1) / 2.0), inv_half_h = 1.f / (float)((H - 1) / 2.0);
  float w1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) w1[j] = s_w1[c0 + j];
  const float* hyp = P.per_pixel ? P.hyp + (long long)b * D * HW + (long long)y * W + x : P.hyp + (long long)b * D;
  const long long hyp_stride = P.per_pixel ? HW : 1;
  VT* vol = reinterpret_cast<VT*>(P.vol) + g8_offset(b, q, 0, y, x, C / 8, D, H, W);
  const long long vol_stride = HW * 8;
  const long long pix = (long long)y * W + x;
  const float inv_nsrc = 1.f / (float)n_src;
  float gref[8], gw1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gref[j] = gw1[j] = 0.f;

  for (int d = 0; d < D; ++d) {
    const float dep = __ldg(hyp + d * hyp_stride);
    F8 gv;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j] = 0.f; gv.v[j] = 0.f; }
    if (OP == OP_GWT || OP == OP_MERGED_BWD) gv = load8(vol + d * vol_stride);
    for (int v = 0; v < n_src; ++v) {
      float ix, iy, w[4], wv[8], df[8], e[8];
      project_b(s_rt + v * 12, fx, fy, dep, inv_half_w, inv_half_h, fw, fh, ix, iy);
      const int off = footprint_b(ix, iy, H, W, C, w);
      blend8(P.src[v] + (long long)b * img_stride + off + c0, W, C, w, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) { df[j] = rf.v[j] - wv[j]; e[j] = df[j] * df[j]; }
      const long long sidx = ((long long)v * P.B + b) * DHW + (long long)d * HW + pix;   // scalar-volume index
      if (OP == OP_SCORE_FWD) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(w1[j], e[j], s);
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (q == 0 && live) P.s_vol[sidx] = s;
      } else if (OP == OP_WEIGHTED_FWD) {
        const float wt1 = __ldg(P.wt_vol + sidx) + 1.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt1, e[j], acc[j]);
      } else if (OP == OP_GWT) {
        // d loss / d wt_v = sum_c g_vol[c] e[c] / n_src: the only thing the scalar chain's backward needs
        float gwt = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) gwt = fmaf(gv.v[j] * inv_nsrc, e[j], gwt);
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) gwt += __shfl_xor_sync(0xffffffffu, gwt, o);
        if (q == 0 && live) P.s_vol[sidx] = gwt;
      } else {
        // merged backward: both paths into e at once, through the aggregate (weight held fixed) and through the score
        float ge[8], g[8];
        const float wt1 = __ldg(P.wt_vol + sidx) + 1.f, gs = __ldg(P.s_vol + sidx);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ge[j] = fmaf(gv.v[j] * inv_nsrc, wt1, gs * w1[j]);
          if (live) gw1[j] = fmaf(gs, e[j], gw1[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gdf = 2.f * df[j] * ge[j];
          gref[j] += gdf;
          g[j] = -gdf;
        }
        if (live) scatter8(P.g_src[v] + (long long)b * img_stride + off + c0, W, C, w, g);
      }
    }
    if (OP == OP_WEIGHTED_FWD && live) {
      F8 r;
#pragma unroll
      for (int j = 0; j < 8; ++j) r.v[j] = acc[j] * inv_nsrc;
      store8(vol + d * vol_stride, r);
    }
  }
  if (OP == OP_MERGED_BWD) {
    if (live) {
      float* gr = P.g_ref + (long long)b * img_stride + pix * C + c0;   // this thread is the only writer of its 8 channels
      F8 r = load8(gr);
#pragma unroll
      for (int j = 0; j < 8; ++j) r.v[j] += gref[j];
      store8(gr, r);
    }
  }
  if (OP == OP_MERGED_BWD && P.g_w1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_gw[c0 + j], gw1[j]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x)
      if (s_gw[i] != 0.f) atomicAdd(P.g_w1 + i, s_gw[i]);
  }
}

template <int C, int OP>
static int launch_train(const WarpTrainParams& P, int vdtype, cudaStream_t st) {
  constexpr int TW = 32 / (C / 8), TH = 4;
  dim3 grid((P.W + TW - 1) / TW, (P.H + TH - 1) / TH, P.B);
  if (vdtype == DAMVS_BF16) warp_train_kernel<C, OP, __nv_bfloat16><<<grid, 128, 0, st>>>(P);
  else warp_train_kernel<C, OP, float><<<grid, 128, 0, st>>>(P);
  DAMVS_LAUNCH_OK("warp_train kernel");
  return DAMVS_OK;
}

template <int OP>
static int dispatch_train(const WarpTrainParams& P, int C, int vdtype, cudaStream_t st) {
  switch (C) {
    case 8: return launch_train<8, OP>(P, vdtype, st);
    case 16: return launch_train<16, OP>(P, vdtype, st);
    case 32: return launch_train<32, OP>(P, vdtype, st);
    case 64: return launch_train<64, OP>(P, vdtype, st);
    default: return set_error(DAMVS_ERR_UNSUPPORTED, "warp_train: C=%d not in {8,16,32,64}", C);
  }
}

static int fill_common(WarpTrainParams& P, const float* ref, const float* const* src, float* const* g_src, int n_src,
                       const float* rot_trans, const float* hyp, int B, int D, int H, int W, int per_pixel) {
  DAMVS_REQUIRE(ref && src && rot_trans && hyp, "warp_train: null pointer");
  DAMVS_REQUIRE(n_src >= 1 && n_src <= kMaxSrcB, "warp_train: n_src=%d outside [1,%d]", n_src, kMaxSrcB);
  DAMVS_REQUIRE(B > 0 && B <= 65535 && D > 0 && H > 1 && W > 1, "warp_train: bad shape");
  DAMVS_REQUIRE(aligned16(ref), "warp_train: ref must be 16-byte aligned");
  P = WarpTrainParams{};
  P.ref = ref;
  for (int v = 0; v < n_src; ++v) {
    DAMVS_REQUIRE(src[v] && aligned16(src[v]), "warp_train: src[%d] null or misaligned", v);
    P.src[v] = src[v];
    if (g_src) {
      DAMVS_REQUIRE(g_src[v] && aligned16(g_src[v]), "warp_train: g_src[%d] null or misaligned", v);
      P.g_src[v] = g_src[v];
    }
  }
  P.rot_trans = rot_trans; P.hyp = hyp; P.B = B; P.n_src = n_src; P.D = D; P.H = H; P.W = W; P.per_pixel = per_pixel;
  return DAMVS_OK;
}

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_warp_score_fwd(const float* ref_nhwc, const float* const* src_nhwc, int n_src, const float* rot_trans,
                                    const float* depth_hyp, const float* w1, float* s_vol, int B, int C, int D, int H, int W,
                                    int per_pixel_hyp, void* stream) {
  WarpTrainParams P;
  int rc = fill_common(P, ref_nhwc, src_nhwc, nullptr, n_src, rot_trans, depth_hyp, B, D, H, W, per_pixel_hyp);
  if (rc) return rc;
  DAMVS_REQUIRE(w1 && s_vol, "warp_score_fwd: null pointer");
  P.w1 = w1; P.s_vol = s_vol;
  return dispatch_train<OP_SCORE_FWD>(P, C, DAMVS_F32, (cudaStream_t)stream);
}

extern "C" int damvs_warp_weighted_fwd(const float* ref_nhwc, const float* const* src_nhwc, int n_src, const float* rot_trans,
                                       const float* depth_hyp, const float* wt_vol, void* out_vol, int B, int C, int D, int H,
                                       int W, int per_pixel_hyp, int out_dtype, void* stream) {
  WarpTrainParams P;
  int rc = fill_common(P, ref_nhwc, src_nhwc, nullptr, n_src, rot_trans, depth_hyp, B, D, H, W, per_pixel_hyp);
  if (rc) return rc;
  DAMVS_REQUIRE(wt_vol && out_vol && aligned16(out_vol), "warp_weighted_fwd: null or misaligned pointer");
  DAMVS_REQUIRE(out_dtype == DAMVS_F32 || out_dtype == DAMVS_BF16, "warp_weighted_fwd: bad out_dtype");
  P.wt_vol = wt_vol; P.vol = out_vol;
  return dispatch_train<OP_WEIGHTED_FWD>(P, C, out_dtype, (cudaStream_t)stream);
}

extern "C" int damvs_warp_gwt(const float* ref_nhwc, const float* const* src_nhwc, int n_src, const float* rot_trans, const float* depth_hyp,
                              const void* g_vol, int g_dtype, float* g_wt_vol, int B, int C, int D, int H, int W, int per_pixel_hyp,
                              void* stream) {
  WarpTrainParams P;
  int rc = fill_common(P, ref_nhwc, src_nhwc, nullptr, n_src, rot_trans, depth_hyp, B, D, H, W, per_pixel_hyp);
  if (rc) return rc;
  DAMVS_REQUIRE(g_vol && g_wt_vol && aligned16(g_vol), "warp_gwt: null or misaligned pointer");
  DAMVS_REQUIRE(g_dtype == DAMVS_F32 || g_dtype == DAMVS_BF16, "warp_gwt: bad g_dtype");
  P.vol = const_cast<void*>(g_vol); P.s_vol = g_wt_vol;
  return dispatch_train<OP_GWT>(P, C, g_dtype, (cudaStream_t)stream);
}

extern "C" int damvs_warp_merged_bwd(const float* ref_nhwc, const float* const* src_nhwc, int n_src, const float* rot_trans,
                                     const float* depth_hyp, const float* w1, const float* wt_vol, const float* g_s_vol, const void* g_vol,
                                     int g_dtype, float* g_ref, float* const* g_src, float* g_w1, int B, int C, int D, int H, int W,
                                     int per_pixel_hyp, void* stream) {
  WarpTrainParams P;
  int rc = fill_common(P, ref_nhwc, src_nhwc, g_src, n_src, rot_trans, depth_hyp, B, D, H, W, per_pixel_hyp);
  if (rc) return rc;
  DAMVS_REQUIRE(w1 && wt_vol && g_s_vol && g_vol && g_ref && g_src && aligned16(g_ref) && aligned16(g_vol),
                "warp_merged_bwd: null or misaligned pointer");
  DAMVS_REQUIRE(g_dtype == DAMVS_F32 || g_dtype == DAMVS_BF16, "warp_merged_bwd: bad g_dtype");
  P.w1 = w1; P.wt_vol = wt_vol; P.s_vol = const_cast<float*>(g_s_vol); P.vol = const_cast<void*>(g_vol); P.g_ref = g_ref; P.g_w1 = g_w1;
  return dispatch_train<OP_MERGED_BWD>(P, C, g_dtype, (cudaStream_t)stream);
}
