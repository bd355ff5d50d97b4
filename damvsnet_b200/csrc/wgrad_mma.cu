// Convolution weight gradients on the tensor cores (training path, bf16 volumes).
//
// dW[co][ci][tap] = sum over voxels of g_y[o][co] * x[i(o, tap)][ci]  -- a GEMM whose reduction dimension is
// the voxel index (reference: autograd of nn.Conv3d / nn.ConvTranspose3d, models/module.py:139, 182).  The
// G8 layout stores a voxel's 8 channels as one 16-byte row, i.e. both operands are "voxel-major"; ldmatrix.trans
// turns eight such rows into the fragment layout mma.sync.m16n8k16 wants (A: 16 ci x 16 voxels, B: 16 voxels x
// 8 co), so no transpose pass and no im2col are needed: a tap is a shifted row address into a shared-memory
// tile with halo.  Voxel index K is huge and M, N are tiny (<= 32 x 8 per CTA), which is the wrong shape for
// tcgen05 (M >= 64) and a good one for warp-level MMA.
//
// CTA = 9 warps, warp w owns (kd, kh) = (w / 3, w % 3) and loops kw; a CTA owns one (32-ci, 8-co) channel block
// (grid.y) and walks "tile columns" (TY x 32 base voxels) through all depth planes, accumulating in registers;
// partial sums are added to dW with fp32 atomics once per CTA.  Base voxels are output voxels for Conv3d
// (stride 1/2: x rows at stride*o - 1 + tap) and input voxels for ConvTranspose3d (g_y rows at 2*i - 1 + tap).
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace damvs {

enum { WG_S1 = 0, WG_S2 = 1, WG_T = 2 };

struct WgMmaParams {
  const __nv_bfloat16* x;   // [B][Gin][Din][Hin][Win][8]
  const __nv_bfloat16* gy;  // [B][Gout][Dout][Hout][Wout][8]
  float* dw;
  int B, Cin, Cout, Gin, Gout, Din, Hin, Win, Dout, Hout, Wout;
  int tiles_x, tiles_y, ncols, transposed;
};

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// MODE: WG_S1 / WG_S2 (Conv3d stride 1 / 2), WG_T (ConvTranspose3d k3 s2 p1 op1).  GA = ci groups per CTA (1, 2 or 4).
template <int MODE, int GA>
__global__ void __launch_bounds__(288) conv_wgrad_mma_kernel(const WgMmaParams P) {
  constexpr int TX = 32, TY = MODE == WG_S1 ? 8 : 4;
  constexpr int S = MODE == WG_S1 ? 1 : 2;
  constexpr int SR = S * (TY - 1) + 3, SC = S * (TX - 1) + 3;   // rows / columns of the shifted operand's tile (with halo)
  constexpr int MT = GA >= 2 ? GA / 2 : 1;                       // 16-row m-tiles
  // shifted operand: x (conv) with GA groups, or g_y (transposed) with 1 group; fixed operand: the other one
  constexpr int GS = MODE == WG_T ? 1 : GA, GF = MODE == WG_T ? GA : 1;
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* sS = reinterpret_cast<uint4*>(smem);                    // [GS][3][SR][SC]
  uint4* sF = sS + GS * 3 * SR * SC;                             // [GF][TY][TX]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kd = warp / 3, kh = warp % 3;
  const int nco = (P.Cout + 7) / 8;
  const int cob = blockIdx.y % nco, cib = blockIdx.y / nco;      // co group, ci block (GA groups)
  // base domain (the GEMM's K index): output voxels for conv, input voxels for transposed
  const int Db = MODE == WG_T ? P.Din : P.Dout, Hb = MODE == WG_T ? P.Hin : P.Hout, Wb = MODE == WG_T ? P.Win : P.Wout;
  // shifted-operand volume
  const __nv_bfloat16* vs = MODE == WG_T ? P.gy : P.x;
  const __nv_bfloat16* vf = MODE == WG_T ? P.x : P.gy;
  const int Ds = MODE == WG_T ? P.Dout : P.Din, Hs = MODE == WG_T ? P.Hout : P.Hin, Ws = MODE == WG_T ? P.Wout : P.Win;
  const int Gs_tot = MODE == WG_T ? P.Gout : P.Gin, Gf_tot = MODE == WG_T ? P.Gin : P.Gout;
  const int gs0 = MODE == WG_T ? cob : cib * GA, gf0 = MODE == WG_T ? cib * GA : cob;

  float acc[3][MT][4];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][m][j] = 0.f;

  const uint32_t sS_addr = (uint32_t)__cvta_generic_to_shared(sS), sF_addr = (uint32_t)__cvta_generic_to_shared(sF);
  // ldmatrix row this lane addresses: matrix mi = lane / 8, row r = lane % 8
  const int mi = lane >> 3, r8 = lane & 7;

  for (int col = blockIdx.x; col < P.ncols; col += gridDim.x) {
    const int b = col / (P.tiles_x * P.tiles_y);
    const int y0 = ((col / P.tiles_x) % P.tiles_y) * TY, x0 = (col % P.tiles_x) * TX;
    for (int z = 0; z < Db; ++z) {
      __syncthreads();   // previous iteration's fragments are consumed
      // ---- shifted operand tile: planes S*z - 1 .. + 2, rows S*y0 - 1 .., columns S*x0 - 1 .. (zero outside the volume)
      for (int i = threadIdx.x; i < GS * 3 * SR * SC; i += blockDim.x) {
        const int c = i % SC, rr = (i / SC) % SR, p = (i / (SC * SR)) % 3, g = i / (SC * SR * 3);
        const int zz = S * z - 1 + p, yy = S * y0 - 1 + rr, xx = S * x0 - 1 + c;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (zz >= 0 && zz < Ds && yy >= 0 && yy < Hs && xx >= 0 && xx < Ws && gs0 + g < Gs_tot)
          v = __ldg(reinterpret_cast<const uint4*>(vs + g8_offset(b, gs0 + g, zz, yy, xx, Gs_tot, Ds, Hs, Ws)));
        sS[i] = v;
      }
      // ---- fixed operand tile: plane z, rows y0 .., columns x0 .. (zero outside the base domain)
      for (int i = threadIdx.x; i < GF * TY * TX; i += blockDim.x) {
        const int c = i % TX, rr = (i / TX) % TY, g = i / (TX * TY);
        const int yy = y0 + rr, xx = x0 + c;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (yy < Hb && xx < Wb && gf0 + g < Gf_tot)
          v = __ldg(reinterpret_cast<const uint4*>(vf + g8_offset(b, gf0 + g, z, yy, xx, Gf_tot, Db, Hb, Wb)));
        sF[i] = v;
      }
      __syncthreads();
      // ---- 16-voxel chunks along x: (row ry, half cx)
#pragma unroll 2
      for (int ch = 0; ch < TY * 2; ++ch) {
        const int ry = ch >> 1, cx = (ch & 1) * 16;
        // fixed-operand fragment (the same for the three kw taps)
        uint32_t f[MT][4];
        if (MODE == WG_T) {
          // A = x: matrices (vox 0-7, g), (vox 0-7, g+1), (vox 8-15, g), (vox 8-15, g+1)
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            if (GA >= 2) {
              const int vox = cx + r8 + 8 * (mi >> 1), g = 2 * m + (mi & 1);
              ldsm_x4_t(sF_addr + (((g * TY + ry) * TX + vox) << 4), f[m][0], f[m][1], f[m][2], f[m][3]);
            } else {
              const int vox = cx + r8 + 8 * (mi & 1);
              ldsm_x2_t(sF_addr + (((ry) * TX + vox) << 4), f[m][0], f[m][2]);
              f[m][1] = f[m][3] = 0u;
            }
          }
        } else {
          const int vox = cx + r8 + 8 * (mi & 1);
          ldsm_x2_t(sF_addr + ((ry * TX + vox) << 4), f[0][0], f[0][1]);
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          if (MODE == WG_T) {
            // B = g_y rows at 2*v - 1 + tap (tile coordinates: 2*v + tap)
            uint32_t b0, b1;
            const int vox = cx + r8 + 8 * (mi & 1);
            ldsm_x2_t(sS_addr + ((((kd) * SR + (S * ry + kh)) * SC + (S * vox + kw)) << 4), b0, b1);
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_bf16_16816(acc[kw][m], f[m][0], f[m][1], f[m][2], f[m][3], b0, b1);
          } else {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              uint32_t a0, a1, a2, a3;
              if (GA >= 2) {
                const int vox = cx + r8 + 8 * (mi >> 1), g = 2 * m + (mi & 1);
                ldsm_x4_t(sS_addr + (((((g * 3 + kd) * SR) + (S * ry + kh)) * SC + (S * vox + kw)) << 4), a0, a1, a2, a3);
              } else {
                const int vox = cx + r8 + 8 * (mi & 1);
                ldsm_x2_t(sS_addr + ((((kd) * SR + (S * ry + kh)) * SC + (S * vox + kw)) << 4), a0, a2);
                a1 = a3 = 0u;
              }
              mma_bf16_16816(acc[kw][m], a0, a1, a2, a3, f[0][0], f[0][1]);
            }
          }
        }
      }
    }
  }
  // ---- D fragment: rows (ci) g, g + 8; columns (co) 2t, 2t + 1
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
    const int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = cib * GA * 8 + m * 16 + g + (j >> 1) * 8, co = cob * 8 + 2 * t + (j & 1);
        const float v = acc[kw][m][j];
        if (ci < P.Cin && co < P.Cout && (GA >= 2 || (j >> 1) == 0) && v != 0.f) {
          float* dst = P.transposed ? P.dw + ((size_t)ci * P.Cout + co) * 27 + tap : P.dw + ((size_t)co * P.Cin + ci) * 27 + tap;
          atomicAdd(dst, v);
        }
      }
  }
}

template <int MODE, int GA>
static int launch_wg(const WgMmaParams& P0, cudaStream_t st) {
  WgMmaParams P = P0;
  constexpr int TX = 32, TY = MODE == WG_S1 ? 8 : 4, S = MODE == WG_S1 ? 1 : 2;
  constexpr int SR = S * (TY - 1) + 3, SC = S * (TX - 1) + 3;
  constexpr int GS = MODE == WG_T ? 1 : GA, GF = MODE == WG_T ? GA : 1;
  const size_t smem = (size_t)(GS * 3 * SR * SC + GF * TY * TX) * 16;
  const int Hb = MODE == WG_T ? P.Hin : P.Hout, Wb = MODE == WG_T ? P.Win : P.Wout;
  P.tiles_x = (Wb + TX - 1) / TX;
  P.tiles_y = (Hb + TY - 1) / TY;
  P.ncols = P.tiles_x * P.tiles_y * P.B;
  auto kern = conv_wgrad_mma_kernel<MODE, GA>;
  DAMVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblk = ((P.Cin / 8 + GA - 1) / GA) * ((P.Cout + 7) / 8);
  const int per_sm = std::max(1, (int)((220 * 1024) / (smem + 1024)));
  const int gx = std::max(1, std::min(P.ncols, std::max(1, 148 * std::min(per_sm, 3) / std::min(nblk, 4))));
  dim3 grid(gx, nblk);
  kern<<<grid, 288, smem, st>>>(P);
  DAMVS_LAUNCH_OK("conv_wgrad_mma kernel");
  return DAMVS_OK;
}

// bf16 x and g_y only; returns DAMVS_ERR_UNSUPPORTED for anything else (the caller falls back to the CUDA-core kernel)
int conv_wgrad_mma_launch(const damvs_conv3d_desc* d, const void* x, const void* g_y, float* dw, cudaStream_t st) {
  if (d->in_dtype != DAMVS_BF16 || d->out_dtype != DAMVS_BF16 || d->Cin % 8) return DAMVS_ERR_UNSUPPORTED;
  WgMmaParams P{};
  P.x = (const __nv_bfloat16*)x; P.gy = (const __nv_bfloat16*)g_y; P.dw = dw;
  P.B = d->B; P.Cin = d->Cin; P.Cout = d->Cout; P.Gin = d->Cin / 8; P.Gout = (d->Cout + 7) / 8;
  P.Din = d->Din; P.Hin = d->Hin; P.Win = d->Win; P.transposed = d->transposed;
  if (d->transposed) { P.Dout = 2 * d->Din; P.Hout = 2 * d->Hin; P.Wout = 2 * d->Win; }
  else { P.Dout = (d->Din - 1) / d->stride + 1; P.Hout = (d->Hin - 1) / d->stride + 1; P.Wout = (d->Win - 1) / d->stride + 1; }
  const int mode = d->transposed ? WG_T : (d->stride == 2 ? WG_S2 : WG_S1);
  const int ga = P.Gin >= 4 ? 4 : (P.Gin >= 2 ? 2 : 1);
#define GO(M, G) if (mode == M && ga == G) return launch_wg<M, G>(P, st)
  GO(WG_S1, 1); GO(WG_S1, 2); GO(WG_S1, 4);
  GO(WG_S2, 1); GO(WG_S2, 2); GO(WG_S2, 4);
  GO(WG_T, 1); GO(WG_T, 2); GO(WG_T, 4);
#undef GO
  return DAMVS_ERR_UNSUPPORTED;
}

}  // namespace damvs
