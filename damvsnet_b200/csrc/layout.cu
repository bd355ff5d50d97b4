// Layout repacks at the boundary of the hot path.
//
// The reference hands over NCHW features and NCDHW volumes; the kernels want
// channels innermost (NHWC features, G8 volumes).  Each repack is a tiled
// transpose through shared memory so that both the read and the write side are
// coalesced; each moves its tensor exactly once (read + write = 2x its bytes).
#include "common.cuh"

namespace damvs {

constexpr int kTilePix = 128;

// in [B][C][HW] -> out [B][HW][C]
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           int C, long long HW) {
  extern __shared__ float tile[];  // [C][kTilePix+1]
  const int b = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * kTilePix;
  const int np = (int)min((long long)kTilePix, HW - p0);
  const float* src = in + (long long)b * C * HW + p0;
  for (int i = threadIdx.x; i < C * kTilePix; i += blockDim.x) {
    int c = i / kTilePix, p = i - c * kTilePix;
    if (p < np) tile[c * (kTilePix + 1) + p] = __ldcs(src + (long long)c * HW + p);
  }
  __syncthreads();
  float* dst = out + ((long long)b * HW + p0) * C;
  for (int i = threadIdx.x; i < np * C; i += blockDim.x) {
    int p = i / C, c = i - p * C;
    dst[i] = tile[c * (kTilePix + 1) + p];
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
  dim3 grid((unsigned)((HW + kTilePix - 1) / kTilePix), B);
  size_t smem = (size_t)C * (kTilePix + 1) * sizeof(float);
  if (smem > 48 * 1024)
    DAMVS_CUDA_OK(cudaFuncSetAttribute(nchw_to_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nchw_to_nhwc_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(in, out, C, HW);
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
  else
    return set_error(DAMVS_ERR_INVALID, "g8_to_ncdhw: bad dtype %d", dtype);
  DAMVS_LAUNCH_OK("g8_to_ncdhw");
  return DAMVS_OK;
}
