// Cross-view photometric loss, fused forward and backward (the training-side neighbour).
//
// Replaces reference cross_view_loss (models/module.py:624-691) and inverse_warping (models/homography.py:7-201), which
// per stage make 2 x (N-1) warps out of ~30 tensor ops each (flat-index gathers, four full-size weight maps, .cuda()
// constants) and back-propagate through all of them.  The structure of the loss makes it cheap to fuse: per source
// view v the reference forms ONE scalar L_v = mean smooth-L1(mask * warp(depth_est), mask * warp(depth_gt)) and
// broadcasts it over the pixels where both warps are valid; per pixel the two smallest valid L_v are summed.  So
//   pass 1 (cross_view_terms):  per pixel and view, both warps in registers -> validity bit + smooth-L1 partial sums
//   pass 2 (cross_view_select): per pixel, the two smallest L_v among the valid views -> how often each view is chosen
//   loss = sum_v count_v / (B H W) * L_v;   backward (cross_view_bwd): d loss / d depth_est, re-warping instead of
//   storing anything; the selection counts are constants, exactly as top-k indices are for autograd.
// Quirks kept (models/homography.py): the source pixel is projected with the REFERENCE intrinsics (:53-57), z + 1e-10
// in the divide, bilinear weights from the CLAMPED x1 / y1, the mask tests y0 <= max_y (:153); floor and clamps carry
// no gradient, so d warp / d x flows through the weights only.
#include <algorithm>

#include "common.cuh"

namespace damvs {

constexpr int kMaxCvViews = 15;

struct CvParams {
  const float* depth_est;                 // [B,H,W]
  const float* depth_gt;                  // [B,H,W]
  const float* img[kMaxCvViews];          // per source view [B,3,H,W] (already resized to the stage resolution)
  const float* cams;                      // [B][n_src][21]: inv(K_ref) [9], proj rows 0-2 [12]
  const float* coeff;                     // bwd: [n_src] = g * w_stage * count_v / (B H W) / (B H W 3)
  uint16_t* maskbits;                     // [B,H,W]
  double* sums;                           // [n_src] smooth-L1 sums (accumulated)
  float* g_depth;                         // [B,H,W]
  int B, n_src, H, W;
};

struct Warp {
  float v[3];       // warped RGB
  float dv[3];      // d warped / d depth
  bool valid;
};

template <bool GRAD>
__device__ __forceinline__ Warp warp_pixel(const float* __restrict__ img, int H, int W, const float* __restrict__ cam, float fx, float fy, float depth) {
  Warp o;
  // cam = inv(K) [x, y, 1] * depth;  p = proj [cam; 1]
  const float r0 = cam[0] * fx + cam[1] * fy + cam[2], r1 = cam[3] * fx + cam[4] * fy + cam[5], r2 = cam[6] * fx + cam[7] * fy + cam[8];
  const float c0 = r0 * depth, c1 = r1 * depth, c2 = r2 * depth;
  const float* P = cam + 9;
  const float p0 = P[0] * c0 + P[1] * c1 + P[2] * c2 + P[3], p1 = P[4] * c0 + P[5] * c1 + P[6] * c2 + P[7], p2 = P[8] * c0 + P[9] * c1 + P[10] * c2 + P[11];
  const float z = p2 + 1e-10f;
  float x = p0 / z, y = p1 / z;
  // _spatial_transformer normalises to [-1,1] and _bilinear_sample un-normalises again (homography.py:118-119, 144-145)
  x = ((x / (float)(W - 1) * 2.f - 1.f) + 1.f) * ((float)W - 1.f) / 2.f;
  y = ((y / (float)(H - 1) * 2.f - 1.f) + 1.f) * ((float)H - 1.f) / 2.f;
  const float fx0 = floorf(x), fy0 = floorf(y);
  int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  o.valid = x0 >= 0 && x1 <= W - 1 && y0 >= 0 && y0 <= H - 1;
  x0 = min(max(x0, 0), W - 1); x1 = min(max(x1, 0), W - 1);
  y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1);
  const float ux = (float)x1 - x, uy = (float)y1 - y;
  const float wa = ux * uy, wb = ux * (1.f - uy), wc = (1.f - ux) * uy, wd = (1.f - ux) * (1.f - uy);
  float dxdd = 0.f, dydd = 0.f;
  if (GRAD) {
    const float q0 = P[0] * r0 + P[1] * r1 + P[2] * r2, q1 = P[4] * r0 + P[5] * r1 + P[6] * r2, q2 = P[8] * r0 + P[9] * r1 + P[10] * r2;
    dxdd = (q0 * z - p0 * q2) / (z * z);
    dydd = (q1 * z - p1 * q2) / (z * z);
  }
  const long long HW = (long long)H * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* im = img + c * HW;
    const float a = __ldg(im + (long long)y0 * W + x0), b = __ldg(im + (long long)y1 * W + x0);
    const float cc = __ldg(im + (long long)y0 * W + x1), d = __ldg(im + (long long)y1 * W + x1);
    o.v[c] = wa * a + wb * b + wc * cc + wd * d;
    if (GRAD) {
      const float dox = -(uy * a + (1.f - uy) * b - uy * cc - (1.f - uy) * d);   // d out / d x  (d ux / d x = -1)
      const float doy = -(ux * a - ux * b + (1.f - ux) * cc - (1.f - ux) * d);   // d out / d y
      o.dv[c] = dox * dxdd + doy * dydd;
    }
  }
  return o;
}

__global__ void __launch_bounds__(128) cross_view_terms_kernel(const __grid_constant__ CvParams P) {
  __shared__ double s_sum[kMaxCvViews];
  if (threadIdx.x < kMaxCvViews) s_sum[threadIdx.x] = 0.0;
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  const long long HW = (long long)P.H * P.W;
  const bool live = x < P.W;
  const long long pix = (long long)b * HW + (long long)y * P.W + (live ? x : 0);
  const float de = __ldg(P.depth_est + pix), dg = __ldg(P.depth_gt + pix);
  uint32_t bits = 0;
  for (int v = 0; v < P.n_src; ++v) {
    const float* cam = P.cams + ((long long)b * P.n_src + v) * 21;
    const float* img = P.img[v] + (long long)b * 3 * HW;
    const Warp we = warp_pixel<false>(img, P.H, P.W, cam, (float)x, (float)y, de);
    const Warp wg = warp_pixel<false>(img, P.H, P.W, cam, (float)x, (float)y, dg);
    const bool m = we.valid && wg.valid && live;
    float part = 0.f;
    if (m) {
      bits |= 1u << v;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = fabsf(we.v[c] - wg.v[c]);
        part += d < 1.f ? 0.5f * d * d : d - 0.5f;      // F.smooth_l1_loss, beta = 1
      }
    }
    // warp reduce, then one shared-memory atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(&s_sum[v], (double)part);
  }
  if (live) P.maskbits[pix] = (uint16_t)bits;
  __syncthreads();
  if (threadIdx.x < P.n_src && s_sum[threadIdx.x] != 0.0) atomicAdd(P.sums + threadIdx.x, s_sum[threadIdx.x]);
}

// per pixel: the two smallest L_v among the valid views (torch.topk(-vol, 2) then the < 1e4 mask, module.py:672-680)
__global__ void __launch_bounds__(256) cross_view_select_kernel(const uint16_t* __restrict__ maskbits, const float* __restrict__ losses, int n_src,
                                                                long long npix, unsigned long long* __restrict__ counts) {
  __shared__ unsigned int s_cnt[kMaxCvViews];
  if (threadIdx.x < kMaxCvViews) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  float L[kMaxCvViews];
  for (int v = 0; v < n_src; ++v) L[v] = __ldg(losses + v);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t bits = maskbits[i];
    int b1 = -1, b2 = -1;
    for (int v = 0; v < n_src; ++v) {
      if (!((bits >> v) & 1u) || !(L[v] < 1e4f)) continue;
      if (b1 < 0 || L[v] < L[b1]) { b2 = b1; b1 = v; }
      else if (b2 < 0 || L[v] < L[b2]) b2 = v;
    }
    if (b1 >= 0) atomicAdd(&s_cnt[b1], 1u);
    if (b2 >= 0) atomicAdd(&s_cnt[b2], 1u);
  }
  __syncthreads();
  if (threadIdx.x < n_src && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(128) cross_view_bwd_kernel(const __grid_constant__ CvParams P) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= P.W) return;
  const long long HW = (long long)P.H * P.W;
  const long long pix = (long long)b * HW + (long long)y * P.W + x;
  const float de = __ldg(P.depth_est + pix), dg = __ldg(P.depth_gt + pix);
  float g = 0.f;
  for (int v = 0; v < P.n_src; ++v) {
    const float cf = __ldg(P.coeff + v);
    if (cf == 0.f) continue;
    const float* cam = P.cams + ((long long)b * P.n_src + v) * 21;
    const float* img = P.img[v] + (long long)b * 3 * HW;
    const Warp we = warp_pixel<true>(img, P.H, P.W, cam, (float)x, (float)y, de);
    if (!we.valid) continue;
    const Warp wg = warp_pixel<false>(img, P.H, P.W, cam, (float)x, (float)y, dg);
    if (!wg.valid) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = we.v[c] - wg.v[c];
      g += cf * fminf(fmaxf(d, -1.f), 1.f) * we.dv[c];    // d smooth_l1 / d warped_est
    }
  }
  P.g_depth[pix] = g;
}

static int fill_cv(CvParams& P, const float* depth_est, const float* depth_gt, const float* const* imgs, const float* cams, int B, int n_src,
                   int H, int W) {
  DAMVS_REQUIRE(depth_est && depth_gt && imgs && cams, "cross_view: null pointer");
  DAMVS_REQUIRE(n_src >= 1 && n_src <= kMaxCvViews, "cross_view: n_src=%d outside [1,%d]", n_src, kMaxCvViews);
  DAMVS_REQUIRE(B > 0 && B <= 65535 && H > 1 && H <= 65535 && W > 1, "cross_view: bad shape");
  P = CvParams{};
  P.depth_est = depth_est; P.depth_gt = depth_gt; P.cams = cams; P.B = B; P.n_src = n_src; P.H = H; P.W = W;
  for (int v = 0; v < n_src; ++v) {
    DAMVS_REQUIRE(imgs[v], "cross_view: img[%d] is null", v);
    P.img[v] = imgs[v];
  }
  return DAMVS_OK;
}

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_cross_view_terms(const float* depth_est, const float* depth_gt, const float* const* view_imgs, const float* cams,
                                      int B, int n_src, int H, int W, uint16_t* maskbits, double* sums, void* stream) {
  CvParams P;
  int rc = fill_cv(P, depth_est, depth_gt, view_imgs, cams, B, n_src, H, W);
  if (rc) return rc;
  DAMVS_REQUIRE(maskbits && sums, "cross_view_terms: null pointer");
  P.maskbits = maskbits; P.sums = sums;
  dim3 grid((W + 127) / 128, H, B);
  cross_view_terms_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P);
  DAMVS_LAUNCH_OK("cross_view_terms kernel");
  return DAMVS_OK;
}

extern "C" int damvs_cross_view_select(const uint16_t* maskbits, const float* losses, int n_src, long long npix, unsigned long long* counts,
                                       void* stream) {
  DAMVS_REQUIRE(maskbits && losses && counts && npix > 0, "cross_view_select: bad arguments");
  DAMVS_REQUIRE(n_src >= 1 && n_src <= kMaxCvViews, "cross_view_select: n_src=%d outside [1,%d]", n_src, kMaxCvViews);
  const int nb = (int)std::min<long long>((npix + 255) / 256, 148 * 8);
  cross_view_select_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(maskbits, losses, n_src, npix, counts);
  DAMVS_LAUNCH_OK("cross_view_select kernel");
  return DAMVS_OK;
}

extern "C" int damvs_cross_view_bwd(const float* depth_est, const float* depth_gt, const float* const* view_imgs, const float* cams,
                                    const float* coeff, int B, int n_src, int H, int W, float* g_depth, void* stream) {
  CvParams P;
  int rc = fill_cv(P, depth_est, depth_gt, view_imgs, cams, B, n_src, H, W);
  if (rc) return rc;
  DAMVS_REQUIRE(coeff && g_depth, "cross_view_bwd: null pointer");
  P.coeff = coeff; P.g_depth = g_depth;
  dim3 grid((W + 127) / 128, H, B);
  cross_view_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P);
  DAMVS_LAUNCH_OK("cross_view_bwd kernel");
  return DAMVS_OK;
}
