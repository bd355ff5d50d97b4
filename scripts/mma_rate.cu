// Microbenchmark: issue rate of tcgen05.mma (kind::f16, M=128, K=16, SS) for several N,
// operand layouts (no-swizzle interleave vs 128B swizzle) and accumulator patterns.
#include <cstdio>
#include <cuda_runtime.h>
#include "../damvsnet_b200/csrc/tc_common.cuh"
using namespace damvs::tc;

__global__ void __launch_bounds__(128) rate_kernel(int N, int layout, int nacc, int iters, int a_stride16, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  if (threadIdx.x < 32) {
    const bool leader = elect_one();
    const uint32_t a16 = smem_u32(smem) >> 4, b16 = (smem_u32(smem) + 96 * 1024) >> 4;
    const uint32_t idesc = idesc_bf16_m128(N);
    uint64_t adesc0, bdesc;
    if (layout == 0) {  // interleave: SBO 128, LBO = plane stride
      adesc0 = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | (a16 | ((4608u >> 4) << 16));
      bdesc = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | (b16 | (((uint32_t)N) << 16));
    } else {            // 128B swizzle, K-major: SBO = 1024
      adesc0 = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (a16 | (1u << 16));
      bdesc = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (b16 | (1u << 16));
    }
    const uint32_t d0 = tmem, d1 = tmem + (nacc > 1 ? N : 0);
    const uint64_t ad0 = adesc0, ad1 = adesc0 + (uint64_t)a_stride16, ad2 = adesc0 + (uint64_t)(2 * a_stride16);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      if (leader) {
        mma_bf16_ss(d0, ad0, bdesc, idesc, 1u); mma_bf16_ss(d1, ad1, bdesc, idesc, 1u);
        mma_bf16_ss(d0, ad2, bdesc, idesc, 1u); mma_bf16_ss(d1, ad0, bdesc, idesc, 1u);
        mma_bf16_ss(d0, ad1, bdesc, idesc, 1u); mma_bf16_ss(d1, ad2, bdesc, idesc, 1u);
        mma_bf16_ss(d0, ad0, bdesc, idesc, 1u); mma_bf16_ss(d1, ad1, bdesc, idesc, 1u);
        mma_bf16_ss(d0, ad2, bdesc, idesc, 1u); mma_bf16_ss(d1, ad0, bdesc, idesc, 1u);
        mma_bf16_ss(d0, ad1, bdesc, idesc, 1u); mma_bf16_ss(d1, ad2, bdesc, idesc, 1u);
        mma_bf16_ss(d0, ad0, bdesc, idesc, 1u); mma_bf16_ss(d1, ad1, bdesc, idesc, 1u);
        mma_bf16_ss(d0, ad2, bdesc, idesc, 1u); mma_bf16_ss(d1, ad0, bdesc, idesc, 1u);
      }
      __syncwarp();
    }
    if (leader) mma_commit(&bar);
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 64;
  printf("layout N nacc a_stride | cycles/MMA issue | cycles/MMA complete\n");
  for (int layout = 0; layout < 2; ++layout)
    for (int N : {16, 48, 96, 192, 256})
      for (int nacc : {1, 2})
        for (int as : {0, 64}) {
          if (nacc * N > 512) continue;
          rate_kernel<<<1, 128, 200 * 1024>>>(N, layout, nacc, iters, as, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2];
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("%d %3d %d %2d | %7.1f | %7.1f  %s\n", layout, N, nacc, as, h[0] / (iters * 16.0), h[1] / (iters * 16.0),
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  // whole chip: 148 CTAs at once
  for (int N : {48, 192}) {
    rate_kernel<<<148, 128, 200 * 1024>>>(N, 0, 2, iters, 64, d);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("148 CTAs layout 0 N %d: %7.1f cycles/MMA complete\n", N, h[1] / (iters * 16.0));
  }
  return 0;
}
