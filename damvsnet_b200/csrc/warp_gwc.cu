// Fused homography warp + GROUP-WISE CORRELATION aggregation (BASELINE.json north star, configs[4] "groups 4-32").
//
// NOT IN THE REFERENCE: wsmtht520/DAMVSNet aggregates by variance or by its adaptive per-view weights only
// (models/cas_mvsnet.py:14, 34-39).  This is the third aggregation mode the north star names, defined as in the
// group-wise-correlation MVS literature (GwcNet / CVP-style cascades):
//
//     cost[g, d, y, x] = 1/(N-1) * sum_v  1/(C/G) * sum_{c in group g}  ref[c, y, x] * warp_v[c, d, y, x]
//
// with warp_v the reference's own homography warp (models/module.py:297-332: (W-1)/2 normalisation,
// align_corners=False sampling, zero padding per tap, no z > 0 mask).  Its oracle is an own restatement
// (oracle/gwc_oracle.py) built on the pinned warp of oracle/damvs_oracle.py.
//
// Mapping: thread = (pixel, 8 consecutive channels); the C/8 lanes of a pixel are adjacent, so a tap of one pixel is one
// contiguous C * sizeof(feature) segment.  Per (depth, view) lane 0 of a pixel builds the footprint (3 FMA + rcp + 2
// FMA, as warp_agg_fast.cu) and broadcasts it with three shuffles; every lane blends its 8 channels (fp32 accumulate),
// multiplies by the reference feature and sums its channels into its group(s).  Groups wider than 8 channels are
// finished with xor-shuffles.  The N x D warped volume never exists; one pass writes the G-channel G8 volume
// (G = 4 is emitted as an 8-channel volume whose channels 4..7 are zero, so that CostRegNet(in_channels=8) consumes it).
#include <cuda_fp16.h>

#include "common.cuh"

namespace damvs {

constexpr int kMaxSrcG = 15;

struct GwcParams {
  const void* ref;
  const void* src[kMaxSrcG];
  const float* rot_trans;  // [n_src][B][12]
  const float* hyp;        // [B][D][H][W] or [B][D]
  void* out;               // G8 [B][max(G,8)/8][D][H][W][8]
  int B, n_src, D, H, W, per_pixel;
};

__device__ __forceinline__ void tap8(const float* p, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void tap8(const __half* p, float (&o)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __half22float2(h[k]);
    o[2 * k] = f.x; o[2 * k + 1] = f.y;
  }
}

__device__ __forceinline__ void put(float* p, float v) { *p = v; }
__device__ __forceinline__ void put(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void put(__half* p, float v) { *reinterpret_cast<unsigned short*>(p) = (unsigned short)(pack_f16x2(v, 0.f) & 0xffffu); }

template <int C, int G, typename FT, typename OutT>
__global__ void __launch_bounds__(128) warp_gwc_kernel(const GwcParams P) {
  constexpr int LPP = C / 8;                       // lanes per pixel
  constexpr int PPW = 32 / LPP;                    // pixels per warp (along x)
  constexpr int GS = C / G;                        // channels per group
  constexpr int NV = GS >= 8 ? 1 : 8 / GS;         // group values a thread produces
  constexpr int LPG = GS >= 8 ? GS / 8 : 1;        // lanes that share one group
  constexpr int GOUT = G < 8 ? 8 : G;              // channels of the emitted volume
  __shared__ float s_rt[kMaxSrcG * 12];
  const int b = blockIdx.z, H = P.H, W = P.W, D = P.D, n_src = P.n_src;
  {
    // ix = u * W/(W-1) - 0.5: the reference's two normalisations folded into rows 0/1 of [rot | trans]
    const float sx = (float)W / (float)(W - 1), sy = (float)H / (float)(H - 1);
    for (int i = threadIdx.x; i < n_src * 12; i += blockDim.x) {
      const int v = i / 12, j = i - v * 12;
      const float s = (j < 3 || j == 9) ? sx : ((j < 6 || j == 10) ? sy : 1.f);
      s_rt[i] = P.rot_trans[((long long)v * P.B + b) * 12 + j] * s;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane % LPP, pw = lane / LPP;
  const int px = blockIdx.x * PPW + pw, py = blockIdx.y * 4 + warp;
  const bool live = px < W && py < H;
  const int x = live ? px : 0, y = live ? py : 0;
  const long long HW = (long long)H * W;
  const FT* ref = reinterpret_cast<const FT*>(P.ref) + ((long long)b * HW + (long long)y * W + x) * C + q * 8;
  float rf[8];
  tap8(ref, rf);
  const float fx = (float)x, fy = (float)y;
  const float* hyp = P.per_pixel ? P.hyp + (long long)b * D * HW + (long long)y * W + x : P.hyp + (long long)b * D;
  const long long hyp_stride = P.per_pixel ? HW : 1;
  const float norm = 1.f / ((float)n_src * (float)GS);
  OutT* out = reinterpret_cast<OutT*>(P.out);
  const int leader = lane - q;                     // lane 0 of this pixel

  for (int d = 0; d < D; ++d) {
    const float dep = __ldg(hyp + d * hyp_stride);
    float acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.f;
    for (int v = 0; v < n_src; ++v) {
      float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
      int off = 0;
      if (q == 0) {
        const float* rt = s_rt + v * 12;
        const float rx = fmaf(rt[0], fx, fmaf(rt[1], fy, rt[2]));
        const float ry = fmaf(rt[3], fx, fmaf(rt[4], fy, rt[5]));
        const float rz = fmaf(rt[6], fx, fmaf(rt[7], fy, rt[8]));
        const float pxs = fmaf(rx, dep, rt[9]), pys = fmaf(ry, dep, rt[10]), pz = fmaf(rz, dep, rt[11]);
        const float inv = 1.f / pz;
        const float ix = fmaf(pxs, inv, -0.5f), iy = fmaf(pys, inv, -0.5f);
        if (ix > -1.f && ix < (float)W && iy > -1.f && iy < (float)H) {   // NaN / inf coordinates sample zero
          const float fx0 = floorf(ix), fy0 = floorf(iy);
          float wr = ix - fx0, wb = iy - fy0, wl = 1.f - wr, wt = 1.f - wb;
          int x0 = (int)fx0, y0 = (int)fy0;
          if (x0 < 0) { x0 = 0; wl = wr; wr = 0.f; } else if (x0 > W - 2) { x0 = W - 2; wr = wl; wl = 0.f; }
          if (y0 < 0) { y0 = 0; wt = wb; wb = 0.f; } else if (y0 > H - 2) { y0 = H - 2; wb = wt; wt = 0.f; }
          w00 = wl * wt; w01 = wr * wt; w10 = wl * wb; w11 = wr * wb;
          off = (y0 * W + x0) * C;
        }
      }
      if (LPP > 1) {
        w00 = __shfl_sync(0xffffffffu, w00, leader);
        w01 = __shfl_sync(0xffffffffu, w01, leader);
        w10 = __shfl_sync(0xffffffffu, w10, leader);
        w11 = __shfl_sync(0xffffffffu, w11, leader);
        off = __shfl_sync(0xffffffffu, off, leader);
      }
      const FT* img = reinterpret_cast<const FT*>(P.src[v]) + (long long)b * HW * C + off + q * 8;
      float t0[8], t1[8], t2[8], t3[8];
      tap8(img, t0);
      tap8(img + C, t1);
      tap8(img + (long long)W * C, t2);
      tap8(img + (long long)W * C + C, t3);
      float pr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pr[k] = rf[k] * fmaf(t3[k], w11, fmaf(t2[k], w10, fmaf(t1[k], w01, t0[k] * w00)));
      if (GS >= 8) {
        float s = ((pr[0] + pr[1]) + (pr[2] + pr[3])) + ((pr[4] + pr[5]) + (pr[6] + pr[7]));
#pragma unroll
        for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        acc[0] += s;
      } else {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < GS; ++j) s += pr[k * GS + j];
          acc[k] += s;
        }
      }
    }
    if (!live) continue;
    if (GS >= 8) {
      if (q % LPG == 0) {
        const int ch = q / LPG;
        put(out + g8_offset(b, ch / 8, d, y, x, GOUT / 8, D, H, W) + (ch % 8), acc[0] * norm);
        if (G < 8) put(out + g8_offset(b, 0, d, y, x, 1, D, H, W) + G + ch, 0.f);   // zero padding channels (G = 4)
      }
    } else {
      const int ch0 = q * NV;
      OutT* o = out + g8_offset(b, ch0 / 8, d, y, x, GOUT / 8, D, H, W) + (ch0 % 8);
      if (NV == 8) {
        F8 r;
#pragma unroll
        for (int k = 0; k < 8; ++k) r.v[k] = acc[k] * norm;
        store8(o, r);
      } else if (NV == 4) {
        store4(o, acc[0] * norm, acc[1] * norm, acc[2] * norm, acc[3] * norm);
      } else {
#pragma unroll
        for (int k = 0; k < NV; ++k) put(o + k, acc[k] * norm);
      }
      if (G < 8) {   // zero padding channels G..7 of the single output group
#pragma unroll
        for (int k = 0; k < NV; ++k) put(o + G + k, 0.f);
      }
    }
  }
}

template <int C, int G, typename FT>
static int launch_gwc_t(const GwcParams& P, int out_dtype, cudaStream_t st) {
  constexpr int PPW = 32 / (C / 8);
  dim3 grid((P.W + PPW - 1) / PPW, (P.H + 3) / 4, P.B);
  if (out_dtype == DAMVS_F32) warp_gwc_kernel<C, G, FT, float><<<grid, 128, 0, st>>>(P);
  else if (out_dtype == DAMVS_F16) warp_gwc_kernel<C, G, FT, __half><<<grid, 128, 0, st>>>(P);
  else warp_gwc_kernel<C, G, FT, __nv_bfloat16><<<grid, 128, 0, st>>>(P);
  DAMVS_LAUNCH_OK("warp_gwc kernel");
  return DAMVS_OK;
}

template <int C, int G>
static int launch_gwc(const GwcParams& P, int feat_dtype, int out_dtype, cudaStream_t st) {
  return feat_dtype == DAMVS_F16 ? launch_gwc_t<C, G, __half>(P, out_dtype, st) : launch_gwc_t<C, G, float>(P, out_dtype, st);
}

// the tuned fp16-feature kernel with a group-wise epilogue (warp_agg_fast.cu)
bool warp_gwc_fast_supported(int C, int G, int D, int n_src, int feat_dtype, int out_dtype);
int warp_gwc_fast_launch(const void* ref, const void* const* src, int n_src, const float* rot_trans, const float* hyp, void* out, int B, int C,
                         int G, int D, int H, int W, int per_pixel, int out_dtype, cudaStream_t st);

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_warp_gwc_fwd(const void* ref_nhwc, const void* const* src_nhwc, int n_src, const float* rot_trans,
                                  const float* depth_hyp, void* out_vol, int B, int C, int G, int D, int H, int W,
                                  int per_pixel_hyp, int feat_dtype, int out_dtype, void* stream) {
  DAMVS_REQUIRE(ref_nhwc && src_nhwc && rot_trans && depth_hyp && out_vol, "warp_gwc: null pointer");
  DAMVS_REQUIRE(n_src >= 1 && n_src <= kMaxSrcG, "warp_gwc: n_src=%d outside [1,%d]", n_src, kMaxSrcG);
  DAMVS_REQUIRE(B > 0 && B <= 65535 && D > 0 && H > 1 && W > 1, "warp_gwc: bad shape B=%d D=%d H=%d W=%d (H, W >= 2)", B, D, H, W);
  DAMVS_REQUIRE((long long)H * W * C < (1ll << 31), "warp_gwc: feature map too large for 32-bit tap offsets");
  DAMVS_REQUIRE(feat_dtype == DAMVS_F32 || feat_dtype == DAMVS_F16, "warp_gwc: features must be fp32 or fp16 NHWC (feat_dtype %d)", feat_dtype);
  DAMVS_REQUIRE(out_dtype == DAMVS_F32 || out_dtype == DAMVS_BF16 || out_dtype == DAMVS_F16, "warp_gwc: bad out_dtype %d", out_dtype);
  DAMVS_REQUIRE(aligned16(ref_nhwc) && aligned16(out_vol), "warp_gwc: ref and out must be 16-byte aligned");
  for (int v = 0; v < n_src; ++v) DAMVS_REQUIRE(src_nhwc[v] && aligned16(src_nhwc[v]), "warp_gwc: src[%d] null or not 16-byte aligned", v);
  if ((long long)H * W * C * 2 < (1ll << 31) && warp_gwc_fast_supported(C, G, D, n_src, feat_dtype, out_dtype) &&
      (C == 8 || C == 16 || C == 32) && (G == 4 || G == 8 || G == 16 || G == 32) && G <= C)
    return warp_gwc_fast_launch(ref_nhwc, src_nhwc, n_src, rot_trans, depth_hyp, out_vol, B, C, G, D, H, W, per_pixel_hyp, out_dtype,
                                (cudaStream_t)stream);
  GwcParams P;
  P.ref = ref_nhwc;
  for (int v = 0; v < kMaxSrcG; ++v) P.src[v] = v < n_src ? src_nhwc[v] : nullptr;
  for (int v = 0; v < n_src; ++v) DAMVS_REQUIRE(src_nhwc[v] && aligned16(src_nhwc[v]), "warp_gwc: src[%d] null or not 16-byte aligned", v);
  P.rot_trans = rot_trans; P.hyp = depth_hyp; P.out = out_vol;
  P.B = B; P.n_src = n_src; P.D = D; P.H = H; P.W = W; P.per_pixel = per_pixel_hyp;
  cudaStream_t st = (cudaStream_t)stream;
#define GO(CC, GG) if (C == CC && G == GG) return launch_gwc<CC, GG>(P, feat_dtype, out_dtype, st)
  GO(8, 4); GO(8, 8);
  GO(16, 4); GO(16, 8); GO(16, 16);
  GO(32, 4); GO(32, 8); GO(32, 16); GO(32, 32);
#undef GO
  return set_error(DAMVS_ERR_UNSUPPORTED, "warp_gwc: C=%d, G=%d not supported (C in {8,16,32}, G in {4,8,16,32}, G <= C)", C, G);
}
