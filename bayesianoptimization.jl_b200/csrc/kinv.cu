// kinv.cu -- K7a: Sigma^-1 from the resident factor, for the gradient of the log marginal likelihood
// (EXT GaussianProcesses.jl `update_target_and_dtarget!`, reached from the closure of optimizemodel!, reference
// src/models/gp.jl:59-64; SURVEY App. A "dmll": the traces tr(Sigma^-1 dK/dtheta) need every entry of Sigma^-1).
//
// Sigma^-1 = W^T W with W = L^-1.  Column-by-column solves against I leave most of the chip idle (the first column tile is the
// whole critical path), so W is built bottom-up by RECURSIVE BLOCK INVERSION -- at level l adjacent diagonal blocks of 2^l
// panels are merged,  W21 = -W22 (L21 W11)  -- which turns the triangular inversion into a handful of large tile GEMMs, and
// Sigma^-1 is one SYRK-shaped launch over W^T.  Every product is the TMA + DMMA tile GEMM of the factorisation (tilegemm.cuh):
// operands k-major, structural zeros skipped through the k range of each tile.
//   buffers: Ki  [cap][cap]  W (row-major) while it is being built, then Sigma^-1 (symmetric, both triangles)
//            WT  [cap][cap]  W^T (row n = column n of W: k-major for the products that contract over rows of W)
//            TT  [cap][cap]  (L21 W11)^T of the current level
#include <algorithm>
#include "tilegemm.cuh"
#include "handle.h"

namespace b200bo {

struct KinvMaps { CUtensorMap L128, Ki64, WT128, WT64, TT128; };

// the 128 x 128 diagonal blocks of W and W^T are the inverted diagonal blocks K2 left behind
__global__ void kinv_seed_kernel(const double* __restrict__ Linv, const double* __restrict__ LinvT, double* __restrict__ W,
                                 double* __restrict__ WT, int64_t ld) {
  const int b = blockIdx.y;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < NB * NB; e += gridDim.x * blockDim.x) {
    const int r = e >> 7, c = e & 127;
    const int64_t o = ((int64_t)b * NB + r) * ld + (int64_t)b * NB + c;
    W[o] = Linv[(int64_t)b * NB * NB + e];
    WT[o] = LinvT[(int64_t)b * NB * NB + e];
  }
}

// MODE 0:  TT[n][m]  =  sum_k L[m][k] WT[n][k]                m in the lower part [a1,b1), n and k in the upper part [a0,a1), k >= n
// MODE 1:  WT[n][m] = W[m][n] = - sum_k TT[n][k] W[m][k]      k in [a1, m]
// MODE 2:  S[i][j] = S[j][i] = sum_{k >= max(i,j)} WT[i][k] WT[j][k]      (lower tiles only)
// One CTA per 128 (A rows) x 64 (B rows) tile; s = panels per merged half at this level.
template <int MODE>
__global__ void __launch_bounds__(TG_THREADS, 2) kinv_gemm_kernel(double* __restrict__ Ki, double* __restrict__ WT, double* __restrict__ TT,
                                                                  int64_t ld, int s, int nblk, const __grid_constant__ KinvMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  int arow, brow, k0, nch;
  const CUtensorMap *mA, *mB;
  if (MODE == 2) {
    int t = blockIdx.x, ti = 0;
    while (t >= 2 * ti + 2) { t -= 2 * ti + 2; ++ti; }           // tile row ti holds 2 ti + 2 lower tiles
    arow = ti * NB; brow = t * TG_BN;
    k0 = arow > brow ? arow : brow;
    nch = (nblk * NB - k0) / KC;
    mA = &maps.WT128; mB = &maps.WT64;
  } else {
    const int p = blockIdx.z;                                    // pair of adjacent diagonal blocks of s panels each
    const int a0 = 2 * p * s, a1 = a0 + s, b1 = (a1 + s < nblk) ? a1 + s : nblk;
    if (a1 >= nblk) return;
    if (MODE == 0) {
      const int tm = blockIdx.y, tn = blockIdx.x;                // A rows: m (128-row tiles of the lower part), B rows: n (64-row tiles)
      if (tm >= b1 - a1) return;
      arow = (a1 + tm) * NB; brow = a0 * NB + tn * TG_BN;
      k0 = brow; nch = (a1 * NB - k0) / KC;
      mA = &maps.L128; mB = &maps.WT64;
    } else {
      const int tn = blockIdx.y, tm = blockIdx.x;                // A rows: n (128-row tiles of the upper part), B rows: m (64-row tiles)
      if (tm >= 2 * (b1 - a1)) return;
      arow = (a0 + tn) * NB; brow = a1 * NB + tm * TG_BN;
      k0 = a1 * NB; nch = (brow + TG_BN - k0) / KC;
      mA = &maps.TT128; mB = &maps.Ki64;
    }
  }
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  tile_gemm(acc, mA, k0, arow, mB, k0, brow, nch, smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3, rg = rho(g);
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t ra = arow + wm * 32 + mt * 8 + rg, rb = brow + wn * 32 + nt * 8 + q + 4 * e;
        const double v = acc[mt][nt][e];
        if (MODE == 0) {
          TT[rb * ld + ra] = v;                                  // (L21 W11)^T
        } else if (MODE == 1) {
          WT[ra * ld + rb] = -v;
          Ki[rb * ld + ra] = -v;
        } else if (rb <= ra) {                                   // tiles on the diagonal: each element written once, with its mirror
          Ki[ra * ld + rb] = v;
          Ki[rb * ld + ra] = v;
        }
      }
}

cudaError_t make_map2d(CUtensorMap* m, double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

static cudaError_t kinv_buffers(b200bo_handle_s* h) {
  if (h->dKi) return cudaSuccess;
  const uint64_t cap = (uint64_t)h->cap;
  cudaError_t e = cudaMalloc(&h->dKi, sizeof(double) * cap * cap);
  if (e == cudaSuccess) e = cudaMalloc(&h->dWT, sizeof(double) * cap * cap);
  if (e == cudaSuccess) e = cudaMalloc(&h->dTT, sizeof(double) * cap * cap);
  if (e == cudaSuccess) e = make_map2d(&h->tmKi64, h->dKi, cap, cap, cap, TG_BN);
  if (e == cudaSuccess) e = make_map2d(&h->tmWT128, h->dWT, cap, cap, cap, TG_BM);
  if (e == cudaSuccess) e = make_map2d(&h->tmWT64, h->dWT, cap, cap, cap, TG_BN);
  if (e == cudaSuccess) e = make_map2d(&h->tmTT128, h->dTT, cap, cap, cap, TG_BM);
  if (e != cudaSuccess) {     // all or nothing: the next call must not find a half-allocated set
    cudaFree(h->dKi); cudaFree(h->dWT); cudaFree(h->dTT);
    h->dKi = h->dWT = h->dTT = nullptr;
  }
  return e;
}

static KinvMaps kinv_maps(b200bo_handle_s* h) {
  KinvMaps maps;
  maps.L128 = h->tmL; maps.Ki64 = h->tmKi64; maps.WT128 = h->tmWT128; maps.WT64 = h->tmWT64; maps.TT128 = h->tmTT128;
  return maps;
}

// W = L^-1 (row-major, lower triangle; blocks above the diagonal blocks are NOT written) into h->dKi and W^T into h->dWT
cudaError_t launch_linv(b200bo_handle_s* h) {
  const int nblk = (int)(h->Np / NB);
  if (nblk == 0) return cudaSuccess;
  cudaError_t e = kinv_buffers(h);
  if (e != cudaSuccess) return e;
  const KinvMaps maps = kinv_maps(h);
  cudaFuncSetAttribute(kinv_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM);
  cudaFuncSetAttribute(kinv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM);
  kinv_seed_kernel<<<dim3(16, nblk), 256, 0, h->stream>>>(h->dLinv, h->dLinvT, h->dKi, h->dWT, h->ld);
  h->launches++;
  for (int s = 1; s < nblk; s *= 2) {
    const int npairs = (nblk + 2 * s - 1) / (2 * s);
    kinv_gemm_kernel<0><<<dim3(2 * s, s, npairs), TG_THREADS, TG_SMEM, h->stream>>>(h->dKi, h->dWT, h->dTT, h->ld, s, nblk, maps);
    kinv_gemm_kernel<1><<<dim3(2 * s, s, npairs), TG_THREADS, TG_SMEM, h->stream>>>(h->dKi, h->dWT, h->dTT, h->ld, s, nblk, maps);
    h->launches += 2;
  }
  return cudaGetLastError();
}

// Sigma^-1 = W^T W from h->dWT (launch_linv) into h->dKi (row-major, both triangles; overwrites W)
cudaError_t launch_kinv_syrk(b200bo_handle_s* h) {
  const int nblk = (int)(h->Np / NB);
  if (nblk == 0) return cudaSuccess;
  const KinvMaps maps = kinv_maps(h);
  cudaFuncSetAttribute(kinv_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM);
  kinv_gemm_kernel<2><<<nblk * (nblk + 1), TG_THREADS, TG_SMEM, h->stream>>>(h->dKi, h->dWT, h->dTT, h->ld, 0, nblk, maps);
  h->launches++;
  return cudaGetLastError();
}

// Sigma^-1 into h->dKi (row-major, leading dimension h->ld, both triangles)
cudaError_t launch_kinv(b200bo_handle_s* h) {
  cudaError_t e = launch_linv(h);
  if (e == cudaSuccess) e = launch_kinv_syrk(h);
  return e;
}

}  // namespace b200bo
