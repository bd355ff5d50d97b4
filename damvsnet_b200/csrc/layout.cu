// Layout repacks at the boundary of the hot path.
//
// The reference hands over NCHW features and NCDHW volumes; the kernels want
// channels innermost (NHWC features, G8 volumes).  Each repack is a tiled
// transpose through shared memory so that both the read and the write side are
// coalesced; each moves its tensor exactly once (read + write = 2x its bytes).
#include "common.cuh"

namespace damvs {

constexpr int kTilePix = 128;
constexpr int kTilePitch = kTilePix + 4;  // keeps rows 16-byte aligned for float4 stores

// in [B][C][HW] -> out [B][HW][C].  A CTA of 128 threads walks 128-pixel tiles: float4 loads along the pixel
// axis into a [C][128+4] tile (conflict free), then every thread writes one pixel's C channels as float4s.
__global__ void __launch_bounds__(128) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           int C, long long HW, int tiles_per_cta) {
  extern __shared__ __align__(16) float tile[];  // [C][kTilePitch]
  const int b = blockIdx.y;
  const int t = threadIdx.x;
  const bool vec = (HW % 4 == 0);
  for (int it = 0; it < tiles_per_cta; ++it) {
    const long long p0 = ((long long)blockIdx.x * tiles_per_cta + it) * kTilePix;
    if (p0 >= HW) break;
    const int np = (int)min((long long)kTilePix, HW - p0);
    const float* src = in + (long long)b * C * HW + p0;
    if (vec) {
      for (int i = t; i < C * (kTilePix / 4); i += blockDim.x) {
        const int c = i / (kTilePix / 4), p4 = (i - c * (kTilePix / 4)) * 4;
        if (p4 < np) {
          const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (long long)c * HW + p4));
          *reinterpret_cast<float4*>(tile + c * kTilePitch + p4) = v;
        }
      }
    } else {
      for (int i = t; i < C * kTilePix; i += blockDim.x) {
        const int c = i / kTilePix, p = i - c * kTilePix;
        if (p < np) tile[c * kTilePitch + p] = __ldcs(src + (long long)c * HW + p);
      }
    }
    __syncthreads();
    if (t < np) {
      float* dst = out + ((long long)b * HW + p0 + t) * C;
      if (C % 4 == 0) {
        for (int c = 0; c < C; c += 4)
          *reinterpret_cast<float4*>(dst + c) = make_float4(tile[c * kTilePitch + t], tile[(c + 1) * kTilePitch + t],
                                                            tile[(c + 2) * kTilePitch + t], tile[(c + 3) * kTilePitch + t]);
      } else {
        for (int c = 0; c < C; ++c) dst[c] = tile[c * kTilePitch + t];
      }
    }
    __syncthreads();
  }
}

// in [B][C][V] fp32 -> out [B][C/8][V][8] of T
template <typename T>
__global__ void __launch_bounds__(kTilePix) ncdhw_to_g8_kernel(const float* __restrict__ in, T* __restrict__ out,
                                                               int G, long long V) {
  __shared__ float tile[8][kTilePix + 1];
  const int bg = blockIdx.y;  // b * G + g
  const long long v0 = (long long)blockIdx.x * kTilePix;
  const int nv = (int)min((long long)kTilePix, V - v0);
  const float* src = in + (long long)bg * 8 * V + v0;
  const int t = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (t < nv) tile[j][t] = __ldcs(src + (long long)j * V + t);
  __syncthreads();
  if (t < nv) {
    F8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) r.v[j] = tile[j][t];
    store8(out + ((long long)bg * V + v0 + t) * 8, r);
  }
}

template <typename T>
__global__ void __launch_bounds__(kTilePix) g8_to_ncdhw_kernel(const T* __restrict__ in, float* __restrict__ out,
                                                               int G, long long V) {
  __shared__ float tile[8][kTilePix + 1];
  const int bg = blockIdx.y;
  const long long v0 = (long long)blockIdx.x * kTilePix;
  const int nv = (int)min((long long)kTilePix, V - v0);
  const int t = threadIdx.x;
  if (t < nv) {
    F8 r = load8(in + ((long long)bg * V + v0 + t) * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) tile[j][t] = r.v[j];
  }
  __syncthreads();
  float* dst = out + (long long)bg * 8 * V + v0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (t < nv) dst[(long long)j * V + t] = tile[j][t];
}

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_nchw_to_nhwc_f32(const float* in, float* out, int B, int C, int H, int W, void* stream) {
  DAMVS_REQUIRE(in && out, "nchw_to_nhwc: null pointer");
  DAMVS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && C <= 256, "nchw_to_nhwc: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  long long HW = (long long)H * W;
  const long long tiles = (HW + kTilePix - 1) / kTilePix;
  const int tiles_per_cta = tiles >= 4 * 148 * 8 ? 4 : 1;
  dim3 grid((unsigned)((tiles + tiles_per_cta - 1) / tiles_per_cta), B);
  size_t smem = (size_t)C * kTilePitch * sizeof(float);
  DAMVS_REQUIRE(aligned16(in) && aligned16(out), "nchw_to_nhwc: pointers must be 16-byte aligned");
  if (smem > 48 * 1024)
    DAMVS_CUDA_OK(cudaFuncSetAttribute(nchw_to_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nchw_to_nhwc_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(in, out, C, HW, tiles_per_cta);
  DAMVS_LAUNCH_OK("nchw_to_nhwc");
  return DAMVS_OK;
}

extern "C" int damvs_ncdhw_to_g8(const float* in, void* out, int dtype, int B, int C, int D, int H, int W,
                                 void* stream) {
  DAMVS_REQUIRE(in && out, "ncdhw_to_g8: null pointer");
  DAMVS_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && D > 0 && H > 0 && W > 0, "ncdhw_to_g8: bad shape (C must be a multiple of 8)");
  DAMVS_REQUIRE(aligned16(out), "ncdhw_to_g8: out must be 16-byte aligned");
  long long V = (long long)D * H * W;
  int G = C / 8;
  dim3 grid((unsigned)((V + kTilePix - 1) / kTilePix), B * G);
  if (dtype == DAMVS_F32)
    ncdhw_to_g8_kernel<float><<<grid, kTilePix, 0, (cudaStream_t)stream>>>(in, (float*)out, G, V);
  else if (dtype == DAMVS_BF16)
    ncdhw_to_g8_kernel<__nv_bfloat16><<<grid, kTilePix, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, G, V);
  else if (dtype == DAMVS_F16)
    ncdhw_to_g8_kernel<__half><<<grid, kTilePix, 0, (cudaStream_t)stream>>>(in, (__half*)out, G, V);
  else
    return set_error(DAMVS_ERR_INVALID, "ncdhw_to_g8: bad dtype %d", dtype);
  DAMVS_LAUNCH_OK("ncdhw_to_g8");
  return DAMVS_OK;
}

extern "C" int damvs_g8_to_ncdhw(const void* in, int dtype, float* out, int B, int C, int D, int H, int W,
                                 void* stream) {
  DAMVS_REQUIRE(in && out, "g8_to_ncdhw: null pointer");
  DAMVS_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && D > 0 && H > 0 && W > 0, "g8_to_ncdhw: bad shape (C must be a multiple of 8)");
  DAMVS_REQUIRE(aligned16(in), "g8_to_ncdhw: in must be 16-byte aligned");
  long long V = (long long)D * H * W;
  int G = C / 8;
  dim3 grid((unsigned)((V + kTilePix - 1) / kTilePix), B * G);
  if (dtype == DAMVS_F32)
    g8_to_ncdhw_kernel<float><<<grid, kTilePix, 0, (cudaStream_t)stream>>>((const float*)in, out, G, V);
  else if (dtype == DAMVS_BF16)
    g8_to_ncdhw_kernel<__nv_bfloat16><<<grid, kTilePix, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, out, G, V);
  else if (dtype == DAMVS_F16)
    g8_to_ncdhw_kernel<__half><<<grid, kTilePix, 0, (cudaStream_t)stream>>>((const __half*)in, out, G, V);
  else
    return set_error(DAMVS_ERR_INVALID, "g8_to_ncdhw: bad dtype %d", dtype);
  DAMVS_LAUNCH_OK("g8_to_ncdhw");
  return DAMVS_OK;
}
