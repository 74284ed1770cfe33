// mll.cu -- K7: gradient of the log marginal likelihood (the MAP objective of MAPGPOptimizer).
//
// Replaces EXT GaussianProcesses.jl `update_target_and_dtarget!` called from the closure of optimizemodel!
// (reference src/models/gp.jl:59-64).  With A = alpha alpha' - Sigma^-1 (SURVEY App. A "dmll"):
//   d/dlogNoise = e^{2 logNoise} tr(A);  d/dbeta = sum(alpha);  d/dll_d = 1/2 sum_ij A_ij sf2 psi(r2_ij) z_dij^2;
//   d/dlsigma = sum_ij A_ij K_ij.
// Sigma^-1 comes from launch_kinv (kinv.cu) on the same factor; this pass is one sweep over its lower triangle
// (4 N^2 bytes read; A and K are symmetric) that regenerates K_ij and its derivatives on the fly.  Two-stage fixed-order reduction.
#include "common.cuh"
#include "handle.h"

namespace b200bo {

constexpr int DM_T = 64;        // pair tile edge
constexpr int DM_MAXD = 32;
constexpr int DM_NACC = DM_MAXD + 3;   // [0..D) per-dim, D: sum A.K, D+1: tr(A), D+2: sum(alpha)

template <int FAM>
__global__ void __launch_bounds__(256) dmll_partial_kernel(const double* __restrict__ Kinv, int64_t ld, const double* __restrict__ Z,
                                                           const double* __restrict__ alpha, int N, int D, double sf2,
                                                           double* __restrict__ part) {
  extern __shared__ double sm[];
  double* za = sm;                  // [D][64]
  double* zb = za + D * DM_T;       // [D][64]
  double* aa = zb + D * DM_T;       // [64]
  double* ab = aa + DM_T;           // [64]
  double* red = ab + DM_T;          // [8 warps][DM_NACC]
  // A and K are symmetric: only the tile pairs on and below the diagonal are visited, the strictly lower ones with weight 2
  int bi = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
  while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) ++bi;
  while (bi * (bi + 1) / 2 > (int)blockIdx.x) --bi;
  const int bj = (int)blockIdx.x - bi * (bi + 1) / 2;
  const double wgt = bi == bj ? 1.0 : 2.0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < DM_T * D; e += 256) {
    const int p = e / D, d = e - p * D;
    const int ra = bi * DM_T + p, rb = bj * DM_T + p;
    za[d * DM_T + p] = ra < N ? Z[(int64_t)ra * D + d] : 0.0;
    zb[d * DM_T + p] = rb < N ? Z[(int64_t)rb * D + d] : 0.0;
  }
  if (tid < DM_T) {
    const int ra = bi * DM_T + tid, rb = bj * DM_T + tid;
    aa[tid] = ra < N ? alpha[ra] : 0.0;
    ab[tid] = rb < N ? alpha[rb] : 0.0;
  }
  __syncthreads();
  double acc[DM_NACC];
#pragma unroll
  for (int k = 0; k < DM_NACC; ++k) acc[k] = 0.0;
  const int tx = tid & 15, ty = tid >> 4;
  for (int a = 0; a < 4; ++a) {
    const int li = 4 * ty + a, gi = bi * DM_T + li;
    for (int b = 0; b < 4; ++b) {
      const int lj = 4 * tx + b, gj = bj * DM_T + lj;
      if (gi < N && gj < N) {
        double r2 = 0.0;
        for (int d = 0; d < D; ++d) { const double df = za[d * DM_T + li] - zb[d * DM_T + lj]; r2 = fma(df, df, r2); }
        double phi, psi;
        kern_phi_psi<FAM>(r2, phi, psi);
        const double A = wgt * (aa[li] * ab[lj] - Kinv[(int64_t)gi * ld + gj]);
        const double Ag = A * sf2 * psi;
#pragma unroll
        for (int d = 0; d < DM_MAXD; ++d)
          if (d < D) { const double df = za[d * DM_T + li] - zb[d * DM_T + lj]; acc[d] = fma(Ag, df * df, acc[d]); }
        acc[DM_MAXD] = fma(A, sf2 * phi, acc[DM_MAXD]);
        if (gi == gj) acc[DM_MAXD + 1] += A;
      }
    }
  }
  if (bj == 0 && tx == 0) {   // sum(alpha) once per row
    for (int a = 0; a < 4; ++a) acc[DM_MAXD + 2] += aa[4 * ty + a];
  }
#pragma unroll
  for (int k = 0; k < DM_NACC; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * DM_NACC + k] = v;
  }
  __syncthreads();
  if (tid < DM_NACC) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w * DM_NACC + tid];
    part[(int64_t)blockIdx.x * DM_NACC + tid] = v;
  }
}

// one CTA per accumulator: thread t adds the partials t, t + 256, ... in order, then a fixed shuffle / shared-memory tree (the single
// 35-thread loop over 4096 partials this replaces took 251 us -- a tenth of a gradient evaluation)
__global__ void __launch_bounds__(256) dmll_final_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out) {
  __shared__ double red[8];
  const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double v = 0.0;
  for (int b = tid; b < nblocks; b += 256) v += part[(int64_t)b * DM_NACC + k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[k] = t;
  }
}

// raw sums into h->dscal[8 .. 8 + DM_NACC): [0..32) per-dim, [32] sum A.K, [33] tr(A), [34] sum(alpha)
cudaError_t launch_dmll(b200bo_handle_s* h, int /*mask*/, double* dout, const double* Kinv) {
  const int N = (int)h->N;
  const int T = (N + DM_T - 1) / DM_T;
  const int nblocks = T * (T + 1) / 2;
  if (nblocks == 0) return cudaSuccess;
  const size_t smem = (size_t)(2 * h->D * DM_T + 2 * DM_T + 8 * DM_NACC) * sizeof(double);
  const double sf2 = exp(2.0 * h->hp.lsigma);
#define B200BO_DMLL(F) dmll_partial_kernel<F><<<nblocks, 256, smem, h->stream>>>(Kinv, h->ld, h->dZ, h->dalpha, N, h->D, sf2, h->dpart)
  switch (h->fam) {
    case FAM_SE: B200BO_DMLL(FAM_SE); break;
    case FAM_MAT12: B200BO_DMLL(FAM_MAT12); break;
    case FAM_MAT32: B200BO_DMLL(FAM_MAT32); break;
    default: B200BO_DMLL(FAM_MAT52); break;
  }
#undef B200BO_DMLL
  dmll_final_kernel<<<DM_NACC, 256, 0, h->stream>>>(h->dpart, nblocks, dout);
  h->launches += 2;
  return cudaGetLastError();
}

}  // namespace b200bo
