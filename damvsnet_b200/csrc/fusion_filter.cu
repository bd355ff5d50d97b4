// Geometric-consistency filtering of one reference view against its source views, fused (SURVEY.md section 8f, rank 2:
// the downstream neighbour of the path).
//
// Replaces, per reference view, the numpy + cv2.remap code of reference filter/dypcd.py:98-159
// (reproject_with_depth, check_geometric_consistency) and the accumulation over source views in filter_depth
// (filter/dypcd.py:205-257): the reference builds ~40 full-resolution float64 temporaries per (ref, src) pair on
// one CPU core; here one thread owns one reference pixel, walks the source views, and keeps the nine dynamic masks'
// counters, the consistent-depth sum and the final vote in registers.  HBM traffic: the reference depth and the
// three confidence maps once, four gathered taps of every source depth map, four small outputs.
//
// Arithmetic follows numpy's promotions so that the thresholded masks agree: projections in float64, the casts to
// float32 where the reference has `.astype(np.float32)`, the relative depth test in float32.  cv2.remap
// (INTER_LINEAR, BORDER_CONSTANT 0) is restated from OpenCV's published algorithm (opencv-python 4.x,
// imgproc/imgwarp.cpp remapBilinear): coordinates are rounded to 1/32 pixel, weights are the float32 products of
// the tabulated 1-D weights, taps outside the image contribute the border value 0.
#include "common.cuh"

namespace damvs {

constexpr int kMaxFuseSrc = 10;   // masks exist for i = 2..10 (filter/dypcd.py:151), so dy_range - 1 <= 10

struct FuseParams {
  const float* depth_ref;
  const float* conf[3];            // stage 1, 2, 3 confidence at the reference depth map's resolution
  const float* depth_src[kMaxFuseSrc];
  // per source view, float64: M_rs = E_src inv(E_ref) rows 0-2 [12], K_src [9], inv(K_src) [9], M_sr = E_ref inv(E_src) rows 0-2 [12]
  double cam[kMaxFuseSrc][42];
  double kref_inv[9], kref[9];
  float conf_thr[3];
  double dist_base, rel_diff_base;
  float* depth_avg;
  uint8_t* photo_mask;
  uint8_t* geo_mask;
  uint8_t* final_mask;
  int n_src, H, W;
};

__device__ __forceinline__ float remap_bilinear(const float* __restrict__ img, int H, int W, float x, float y) {
  // cv::remap, CV_32FC1 maps, INTER_LINEAR: sx = cvRound(x * 32) (round half to even), 5 fractional bits
  const int sx = __float2int_rn(x * 32.f), sy = __float2int_rn(y * 32.f);
  const int ix = sx >> 5, iy = sy >> 5;
  const float fx = (float)(sx & 31) * (1.f / 32.f), fy = (float)(sy & 31) * (1.f / 32.f);
  const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
  auto tap = [&](int yy, int xx) -> float {
    return ((unsigned)xx < (unsigned)W && (unsigned)yy < (unsigned)H) ? __ldg(img + (long long)yy * W + xx) : 0.f;
  };
  return tap(iy, ix) * w00 + tap(iy, ix + 1) * w01 + tap(iy + 1, ix) * w10 + tap(iy + 1, ix + 1) * w11;
}

__global__ void __launch_bounds__(128) geo_fuse_kernel(const __grid_constant__ FuseParams P) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= P.W) return;
  const long long pix = (long long)y * P.W + x;
  const float dref = __ldg(P.depth_ref + pix);
  // reference 3-D point: inv(K_ref) [x, y, 1]^T * depth   (float64, filter/dypcd.py:105-106)
  const double dx = (double)x * (double)dref, dy = (double)y * (double)dref, dz = (double)dref;
  const double X = P.kref_inv[0] * dx + P.kref_inv[1] * dy + P.kref_inv[2] * dz;
  const double Y = P.kref_inv[3] * dx + P.kref_inv[4] * dy + P.kref_inv[5] * dz;
  const double Z = P.kref_inv[6] * dx + P.kref_inv[7] * dy + P.kref_inv[8] * dz;
  int geo_sum = 0, sums[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) sums[i] = 0;
  float depth_sum = 0.f;
  for (int v = 0; v < P.n_src; ++v) {
    const double* c = P.cam[v];
    const double sxx = c[0] * X + c[1] * Y + c[2] * Z + c[3];
    const double syy = c[4] * X + c[5] * Y + c[6] * Z + c[7];
    const double szz = c[8] * X + c[9] * Y + c[10] * Z + c[11];
    const double* K = c + 12;
    const double kx = K[0] * sxx + K[1] * syy + K[2] * szz, ky = K[3] * sxx + K[4] * syy + K[5] * szz, kz = K[6] * sxx + K[7] * syy + K[8] * szz;
    const double u = kx / kz, w = ky / kz;                       // xy_src (float64)
    const float samp = remap_bilinear(P.depth_src[v], P.H, P.W, (float)u, (float)w);
    const double* Ki = c + 21;
    const double ax = u * (double)samp, ay = w * (double)samp, az = (double)samp;
    const double px = Ki[0] * ax + Ki[1] * ay + Ki[2] * az, py = Ki[3] * ax + Ki[4] * ay + Ki[5] * az, pz = Ki[6] * ax + Ki[7] * ay + Ki[8] * az;
    const double* M = c + 30;
    const double rx = M[0] * px + M[1] * py + M[2] * pz + M[3];
    const double ry = M[4] * px + M[5] * py + M[6] * pz + M[7];
    const double rz = M[8] * px + M[9] * py + M[10] * pz + M[11];
    float depth_rep = (float)rz;
    double qx = P.kref[0] * rx + P.kref[1] * ry + P.kref[2] * rz, qy = P.kref[3] * rx + P.kref[4] * ry + P.kref[5] * rz,
           qz = P.kref[6] * rx + P.kref[7] * ry + P.kref[8] * rz;
    if (qz == 0.0) qz += 0.00001;
    const float xr = (float)(qx / qz), yr = (float)(qy / qz);
    const double ex = (double)xr - (double)x, ey = (double)yr - (double)y;
    const double dist = sqrt(ex * ex + ey * ey);
    const float rel = fabsf(depth_rep - dref) / dref;            // float32, filter/dypcd.py:145-146
    bool last = false;
#pragma unroll
    for (int i = 2; i <= 10; ++i) {
      const bool m = dist < (double)i * P.dist_base && rel < (float)((double)i * P.rel_diff_base);
      sums[i - 2] += m ? 1 : 0;
      last = m;
    }
    geo_sum += last ? 1 : 0;
    depth_sum += last ? depth_rep : 0.f;                          // depth_reprojected[~mask] = 0
  }
  const int dy_range = P.n_src + 1;
  bool geo = geo_sum >= dy_range;
#pragma unroll
  for (int i = 2; i <= 10; ++i)
    if (i < dy_range) geo = geo || sums[i - 2] >= i;
  const bool photo = __ldg(P.conf[2] + pix) > P.conf_thr[2] && __ldg(P.conf[1] + pix) > P.conf_thr[1] && __ldg(P.conf[0] + pix) > P.conf_thr[0];
  P.depth_avg[pix] = (float)((double)(depth_sum + dref) / (double)(geo_sum + 1));
  P.photo_mask[pix] = photo;
  P.geo_mask[pix] = geo;
  P.final_mask[pix] = photo && geo;
}

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_geo_consistency_fuse(const float* depth_ref, const float* conf1, const float* conf2, const float* conf3,
                                          const float* const* depth_src, const double* cams, int n_src, int H, int W,
                                          float conf_thr1, float conf_thr2, float conf_thr3, double dist_base, double rel_diff_base,
                                          float* depth_avg, uint8_t* photo_mask, uint8_t* geo_mask, uint8_t* final_mask, void* stream) {
  DAMVS_REQUIRE(depth_ref && conf1 && conf2 && conf3 && depth_src && cams && depth_avg && photo_mask && geo_mask && final_mask,
                "geo_consistency_fuse: null pointer");
  DAMVS_REQUIRE(n_src >= 1 && n_src <= kMaxFuseSrc, "geo_consistency_fuse: n_src=%d outside [1,%d]", n_src, kMaxFuseSrc);
  DAMVS_REQUIRE(H > 0 && H <= 65535 && W > 0, "geo_consistency_fuse: bad shape");
  FuseParams P{};
  P.depth_ref = depth_ref; P.conf[0] = conf1; P.conf[1] = conf2; P.conf[2] = conf3;
  // cams: [K_ref 9][inv(K_ref) 9] then per source view 42 doubles (see FuseParams)
  for (int i = 0; i < 9; ++i) { P.kref[i] = cams[i]; P.kref_inv[i] = cams[9 + i]; }
  for (int v = 0; v < n_src; ++v) {
    DAMVS_REQUIRE(depth_src[v], "geo_consistency_fuse: depth_src[%d] is null", v);
    P.depth_src[v] = depth_src[v];
    for (int i = 0; i < 42; ++i) P.cam[v][i] = cams[18 + v * 42 + i];
  }
  P.conf_thr[0] = conf_thr1; P.conf_thr[1] = conf_thr2; P.conf_thr[2] = conf_thr3;
  P.dist_base = dist_base; P.rel_diff_base = rel_diff_base;
  P.depth_avg = depth_avg; P.photo_mask = photo_mask; P.geo_mask = geo_mask; P.final_mask = final_mask;
  P.n_src = n_src; P.H = H; P.W = W;
  dim3 grid((W + 127) / 128, H);
  geo_fuse_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P);
  DAMVS_LAUNCH_OK("geo_fuse kernel");
  return DAMVS_OK;
}
