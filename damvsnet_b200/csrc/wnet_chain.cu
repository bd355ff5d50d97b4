// The scalar tail of AggWeightNetVolume in TRAINING mode, native (reference models/module.py:548-563 applied per source
// view, models/cas_mvsnet.py:71):  s_v -> BatchNorm3d(1) -> ReLU -> 1x1x1 conv (a scalar w2) -> BatchNorm3d(1) -> ReLU -> wt_v
// on the per-view score volumes [n_src][M] (M = B*D*H*W fp32, 1/C of the cost volume), forward and backward.
// Batch statistics make every step a reduction followed by a per-view coefficient update, so the chain is five small
// launches each way instead of ~12 tensor ops per view and direction:
//   fwd:  stats(s) -> finalize(bn1) -> stats(t = w2 relu(a1 s + c1)) -> finalize(bn2) -> wt = relu(a2 t + c2)
//   bwd:  sums(g_pre2, g_pre2 t) -> coeffs(bn2) -> sums(g_pre1, g_pre1 s; g_w2) -> coeffs(bn1) -> g_s
// The running buffers are updated view by view with the module's momentum, as n_src successive forward calls do.
#include <algorithm>

#include "common.cuh"

namespace damvs {

constexpr int kChainMaxViews = 15;

// per-view state, fp32 [n_src][8]: a1, c1, mean1, rstd1, a2, c2, mean2, rstd2
enum { CS_A1 = 0, CS_C1, CS_M1, CS_R1, CS_A2, CS_C2, CS_M2, CS_R2, CS_N };
// per-view backward coefficients, fp32 [n_src][6]: k1, k2, k3 of bn2 then of bn1 (g_in = k1 g_pre + k2 x + k3)
enum { CB_K1B = 0, CB_K2B, CB_K3B, CB_K1A, CB_K2A, CB_K3A, CB_N };

__device__ __forceinline__ void block_sum2(double a, double b, double* dst) {
  __shared__ double sh[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0, y = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { x += sh[0][k]; y += sh[1][k]; }
    atomicAdd(dst, x);
    atomicAdd(dst + 1, y);
  }
  __syncthreads();
}

// STAGE 0: sums of s.  STAGE 1: sums of t = w2 relu(a1 s + c1).
template <int STAGE>
__global__ void __launch_bounds__(256) chain_stats_kernel(const float* __restrict__ s_vol, const float* __restrict__ state, const float* __restrict__ w2p, long long M,
                                                          double* __restrict__ sums) {
  const int v = blockIdx.y;
  const float w2 = __ldg(w2p);
  const float* s = s_vol + (long long)v * M;
  const float a1 = STAGE ? state[v * CS_N + CS_A1] : 0.f, c1 = STAGE ? state[v * CS_N + CS_C1] : 0.f;
  float sa = 0.f, sb = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    float x = __ldg(s + i);
    if (STAGE) x = w2 * fmaxf(fmaf(a1, x, c1), 0.f);
    sa += x;
    sb = fmaf(x, x, sb);
  }
  block_sum2((double)sa, (double)sb, sums + 2 * v);
}

// one thread: per view mean / rstd / folded affine of one BatchNorm, running buffers updated view by view
__global__ void chain_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float* __restrict__ running_mean, float* __restrict__ running_var, double M, float momentum, float eps,
                                      int n_src, int which, float* __restrict__ state) {
  if (threadIdx.x || blockIdx.x) return;
  const double g = gamma[0], b = beta[0];
  double rm = running_mean ? (double)running_mean[0] : 0.0, rv = running_var ? (double)running_var[0] : 0.0;
  for (int v = 0; v < n_src; ++v) {
    const double mean = sums[2 * v] / M;
    double var = sums[2 * v + 1] / M - mean * mean;
    if (var < 0.0) var = 0.0;
    rm = (1.0 - momentum) * rm + momentum * mean;
    rv = (1.0 - momentum) * rv + momentum * var * (M / fmax(M - 1.0, 1.0));
    const double rstd = rsqrt(var + (double)eps), a = g * rstd;
    float* st = state + v * CS_N + (which ? CS_A2 : CS_A1);
    st[0] = (float)a; st[1] = (float)(b - mean * a); st[2] = (float)mean; st[3] = (float)rstd;
  }
  if (running_mean) { running_mean[0] = (float)rm; running_var[0] = (float)rv; }
}

__global__ void __launch_bounds__(256) chain_apply_kernel(const float* __restrict__ s_vol, const float* __restrict__ state, const float* __restrict__ w2p, long long M,
                                                          float* __restrict__ wt_vol) {
  const int v = blockIdx.y;
  const float w2 = __ldg(w2p);
  const float* st = state + v * CS_N;
  const float a1 = st[CS_A1], c1 = st[CS_C1], a2 = st[CS_A2], c2 = st[CS_C2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float t = w2 * fmaxf(fmaf(a1, __ldg(s_vol + (long long)v * M + i), c1), 0.f);
    wt_vol[(long long)v * M + i] = fmaxf(fmaf(a2, t, c2), 0.f);
  }
}

// STAGE 0: sums {g_pre2, g_pre2 t}.  STAGE 1: sums {g_pre1, g_pre1 s} and g_w2 += sum g_t a.  STAGE 2: g_s out.
template <int STAGE>
__global__ void __launch_bounds__(256) chain_bwd_kernel(const float* __restrict__ s_vol, const float* __restrict__ g_wt, const float* __restrict__ state,
                                                        const float* __restrict__ coef, const float* __restrict__ w2p, long long M, double* __restrict__ sums,
                                                        double* __restrict__ g_w2, float* __restrict__ g_s) {
  const int v = blockIdx.y;
  const float w2 = __ldg(w2p);
  const float* st = state + v * CS_N;
  const float a1 = st[CS_A1], c1 = st[CS_C1], a2 = st[CS_A2], c2 = st[CS_C2];
  float k1b = 0.f, k2b = 0.f, k3b = 0.f, k1a = 0.f, k2a = 0.f, k3a = 0.f;
  if (STAGE >= 1) { k1b = coef[v * CB_N + CB_K1B]; k2b = coef[v * CB_N + CB_K2B]; k3b = coef[v * CB_N + CB_K3B]; }
  if (STAGE == 2) { k1a = coef[v * CB_N + CB_K1A]; k2a = coef[v * CB_N + CB_K2A]; k3a = coef[v * CB_N + CB_K3A]; }
  float sa = 0.f, sb = 0.f, sw = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float s = __ldg(s_vol + (long long)v * M + i), gw = __ldg(g_wt + (long long)v * M + i);
    const float pre1 = fmaf(a1, s, c1), a = fmaxf(pre1, 0.f), t = w2 * a, pre2 = fmaf(a2, t, c2);
    const float gp2 = pre2 > 0.f ? gw : 0.f;
    if (STAGE == 0) { sa += gp2; sb = fmaf(gp2, t, sb); continue; }
    const float gt = fmaf(k1b, gp2, fmaf(k2b, t, k3b));     // BatchNorm-2 backward
    const float gp1 = pre1 > 0.f ? gt * w2 : 0.f;
    if (STAGE == 1) { sa += gp1; sb = fmaf(gp1, s, sb); sw = fmaf(gt, a, sw); continue; }
    g_s[(long long)v * M + i] = fmaf(k1a, gp1, fmaf(k2a, s, k3a));   // BatchNorm-1 backward
  }
  if (STAGE <= 1) block_sum2((double)sa, (double)sb, sums + 2 * v);
  if (STAGE == 1) block_sum2((double)sw, 0.0, g_w2);
}

// one thread: BatchNorm backward coefficients per view and the (view-summed) d gamma, d beta
__global__ void chain_bwd_coeffs_kernel(const double* __restrict__ sums, const float* __restrict__ state, double M, int n_src, int which,
                                        float* __restrict__ coef, float* __restrict__ g_gamma, float* __restrict__ g_beta,
                                        const double* __restrict__ gw2_sum, float* __restrict__ g_w2) {
  if (threadIdx.x || blockIdx.x) return;
  if (g_w2) g_w2[0] = (float)gw2_sum[0];
  double gg = 0.0, gb = 0.0;
  for (int v = 0; v < n_src; ++v) {
    const float* st = state + v * CS_N + (which ? CS_A2 : CS_A1);
    const double a = st[0], mean = st[2], rstd = st[3];
    const double sg = sums[2 * v], sgx = sums[2 * v + 1];
    const double dot = rstd * (sgx - mean * sg), m1 = sg / M, m2 = dot / M;
    float* c = coef + v * CB_N + (which ? CB_K1B : CB_K1A);
    c[0] = (float)a; c[1] = (float)(-a * rstd * m2); c[2] = (float)(a * (mean * rstd * m2 - m1));
    gg += dot; gb += sg;
  }
  g_gamma[0] = (float)gg;
  g_beta[0] = (float)gb;
}

static int chain_grid(long long M) { return (int)std::min<long long>((M + 255) / 256, 148 * 4); }

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_wnet_chain_fwd(const float* s_vol, int n_src, long long M, const float* gamma1, const float* beta1, float* running_mean1,
                                    float* running_var1, const float* w2, const float* gamma2, const float* beta2, float* running_mean2,
                                    float* running_var2, float momentum, float eps, double* sums_ws /* [2][n_src][2], zeroed */, float* state,
                                    float* wt_vol, void* stream) {
  DAMVS_REQUIRE(s_vol && gamma1 && beta1 && w2 && gamma2 && beta2 && sums_ws && state && wt_vol, "wnet_chain_fwd: null pointer");
  DAMVS_REQUIRE(n_src >= 1 && n_src <= kChainMaxViews && M > 0, "wnet_chain_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(chain_grid(M), n_src);
  chain_stats_kernel<0><<<grid, 256, 0, st>>>(s_vol, state, w2, M, sums_ws);
  DAMVS_LAUNCH_OK("chain_stats kernel");
  chain_finalize_kernel<<<1, 32, 0, st>>>(sums_ws, gamma1, beta1, running_mean1, running_var1, (double)M, momentum, eps, n_src, 0, state);
  DAMVS_LAUNCH_OK("chain_finalize kernel");
  chain_stats_kernel<1><<<grid, 256, 0, st>>>(s_vol, state, w2, M, sums_ws + 2 * n_src);
  DAMVS_LAUNCH_OK("chain_stats kernel");
  chain_finalize_kernel<<<1, 32, 0, st>>>(sums_ws + 2 * n_src, gamma2, beta2, running_mean2, running_var2, (double)M, momentum, eps, n_src, 1, state);
  DAMVS_LAUNCH_OK("chain_finalize kernel");
  chain_apply_kernel<<<grid, 256, 0, st>>>(s_vol, state, w2, M, wt_vol);
  DAMVS_LAUNCH_OK("chain_apply kernel");
  return DAMVS_OK;
}

extern "C" int damvs_wnet_chain_bwd(const float* s_vol, const float* g_wt, int n_src, long long M, const float* state, const float* w2,
                                    double* sums_ws /* [2][n_src][2] + [2], zeroed */, float* coef_ws /* [n_src][6] */, float* g_s,
                                    float* g_params /* d gamma1, d beta1, d w2, d gamma2, d beta2 */, void* stream) {
  DAMVS_REQUIRE(s_vol && g_wt && state && w2 && sums_ws && coef_ws && g_s && g_params, "wnet_chain_bwd: null pointer");
  DAMVS_REQUIRE(n_src >= 1 && n_src <= kChainMaxViews && M > 0, "wnet_chain_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(chain_grid(M), n_src);
  double* sums2 = sums_ws, *sums1 = sums_ws + 2 * n_src, *gw2 = sums_ws + 4 * n_src;
  chain_bwd_kernel<0><<<grid, 256, 0, st>>>(s_vol, g_wt, state, coef_ws, w2, M, sums2, gw2, g_s);
  DAMVS_LAUNCH_OK("chain_bwd kernel");
  chain_bwd_coeffs_kernel<<<1, 32, 0, st>>>(sums2, state, (double)M, n_src, 1, coef_ws, g_params + 3, g_params + 4, nullptr, nullptr);
  DAMVS_LAUNCH_OK("chain_bwd_coeffs kernel");
  chain_bwd_kernel<1><<<grid, 256, 0, st>>>(s_vol, g_wt, state, coef_ws, w2, M, sums1, gw2, g_s);
  DAMVS_LAUNCH_OK("chain_bwd kernel");
  chain_bwd_coeffs_kernel<<<1, 32, 0, st>>>(sums1, state, (double)M, n_src, 0, coef_ws, g_params + 0, g_params + 1, gw2, g_params + 2);
  DAMVS_LAUNCH_OK("chain_bwd_coeffs kernel");
  chain_bwd_kernel<2><<<grid, 256, 0, st>>>(s_vol, g_wt, state, coef_ws, w2, M, sums1, gw2, g_s);
  DAMVS_LAUNCH_OK("chain_bwd kernel");
  return DAMVS_OK;
}
