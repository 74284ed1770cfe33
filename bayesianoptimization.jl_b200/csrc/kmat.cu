// kmat.cu -- K1: kernel-matrix assembly  Sigma = K(X,X) + (e^{2 logNoise} + eps) I   (FP64, HBM-write bound).
//
// Replaces EXT GaussianProcesses.jl `update_cK!` (cov(kernel, x, x) + noise on the diagonal) reached from
// update!(model, x, y) (reference src/models/gp.jl:11-18).  Math: SURVEY.md Appendix A "Parametrisation"/"Fit".
//
// One CTA per 64x64 tile of the LOWER triangle of tiles; the tile is computed once and stored twice (tile and
// its transpose, the latter through shared memory) so the output is exactly symmetric and every global store is
// written as full 32-byte sectors with 16-byte vector stores (the mirror is the thread's 4x4 block transposed in registers).  Algorithmic bytes: 8 N^2 written + 8 N D read.
#include "common.cuh"
#include "handle.h"

namespace b200bo {

constexpr int KT = 64;   // tile edge

__global__ void scale_inputs_kernel(const double* __restrict__ X, const double* __restrict__ inv_ell, double* __restrict__ Z,
                                    int D, int64_t e0, int64_t e1) {
  const int64_t i = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < e1) Z[i] = X[i] * inv_ell[i % D];
}

template <int FAM>
__global__ void __launch_bounds__(256, 3) kmat_kernel(const double* __restrict__ Z, int N, int Np, int D, double sf2, double noise,
                                                      int pad_identity, double* __restrict__ K, int64_t ld) {
  extern __shared__ double sm[];
  double* za = sm;                 // [D][KT]  rows of the tile (points bi*KT..)
  double* zb = sm + D * KT;        // [D][KT]  cols of the tile (points bj*KT..)
  // linear lower-triangle tile index -> (bi >= bj)
  const int t = blockIdx.x;
  int bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  while (bi * (bi + 1) / 2 > t) --bi;
  const int bj = t - bi * (bi + 1) / 2;
  const int tid = threadIdx.x;
  for (int e = tid; e < KT * D; e += 256) {   // coalesced over the point-major input, transposed into [d][point]
    const int p = e / D, d = e - p * D;
    const int ra = bi * KT + p, rb = bj * KT + p;
    za[d * KT + p] = ra < N ? Z[(int64_t)ra * D + d] : 0.0;
    zb[d * KT + p] = rb < N ? Z[(int64_t)rb * D + d] : 0.0;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double r2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) r2[a][b] = 0.0;
  for (int d = 0; d < D; ++d) {
    const double2 xa0 = *reinterpret_cast<const double2*>(za + d * KT + 4 * ty), xa1 = *reinterpret_cast<const double2*>(za + d * KT + 4 * ty + 2);
    const double2 xb0 = *reinterpret_cast<const double2*>(zb + d * KT + 4 * tx), xb1 = *reinterpret_cast<const double2*>(zb + d * KT + 4 * tx + 2);
    const double xa[4] = {xa0.x, xa0.y, xa1.x, xa1.y}, xb[4] = {xb0.x, xb0.y, xb1.x, xb1.y};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) { const double df = xa[a] - xb[b]; r2[a][b] = fma(df, df, r2[a][b]); }
  }
  const int lim = pad_identity ? Np : N;
  double v[4][4];
  if ((bi + 1) * KT <= N) {
    // ---- interior tile (the common case): no bounds logic at all ----
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) v[a][b] = sf2 * kern_phi<FAM>(r2[a][b]);
    if (bi == bj && ty == tx) {
#pragma unroll
      for (int a = 0; a < 4; ++a) v[a][a] += noise;
    }
    double* p = K + (int64_t)(bi * KT + 4 * ty) * ld + bj * KT + 4 * tx;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      *reinterpret_cast<double2*>(p + (int64_t)a * ld) = make_double2(v[a][0], v[a][1]);
      *reinterpret_cast<double2*>(p + (int64_t)a * ld + 2) = make_double2(v[a][2], v[a][3]);
    }
    if (bi != bj) {   // mirrored tile: the thread's 4 x 4 block transposed in registers, full 32-byte sectors, no staging
      double* m = K + (int64_t)(bj * KT + 4 * tx) * ld + bi * KT + 4 * ty;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        *reinterpret_cast<double2*>(m + (int64_t)b * ld) = make_double2(v[0][b], v[1][b]);
        *reinterpret_cast<double2*>(m + (int64_t)b * ld + 2) = make_double2(v[2][b], v[3][b]);
      }
    }
    return;
  }
  // ---- edge tile: ragged N and the identity padding ----
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = bi * KT + 4 * ty + a;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = bj * KT + 4 * tx + b;
      double val;
      if (gi < N && gj < N) {
        val = sf2 * kern_phi<FAM>(r2[a][b]);
        if (gi == gj) val += noise;
      } else {
        val = (pad_identity && gi == gj) ? 1.0 : 0.0;
      }
      v[a][b] = val;
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = bi * KT + 4 * ty + a, gj = bj * KT + 4 * tx;
    if (gi < lim) {
      double* p = K + (int64_t)gi * ld + gj;
      if (gj + 3 < lim) {
        *reinterpret_cast<double2*>(p) = make_double2(v[a][0], v[a][1]);
        *reinterpret_cast<double2*>(p + 2) = make_double2(v[a][2], v[a][3]);
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b) if (gj + b < lim) p[b] = v[a][b];
      }
    }
  }
  if (bi != bj) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gi = bj * KT + 4 * tx + b, gj = bi * KT + 4 * ty;
      if (gi < lim) {
        double* p = K + (int64_t)gi * ld + gj;
        if (gj + 3 < lim) {
          *reinterpret_cast<double2*>(p) = make_double2(v[0][b], v[1][b]);
          *reinterpret_cast<double2*>(p + 2) = make_double2(v[2][b], v[3][b]);
        } else {
#pragma unroll
          for (int a = 0; a < 4; ++a) if (gj + a < lim) p[a] = v[a][b];
        }
      }
    }
  }
}

cudaError_t launch_scale_inputs(b200bo_handle_s* h, int64_t n0, int64_t n1) {
  const int64_t e0 = n0 * h->D, e1 = n1 * h->D;
  if (e1 <= e0) return cudaSuccess;
  const int blocks = (int)((e1 - e0 + 255) / 256);
  scale_inputs_kernel<<<blocks, 256, 0, h->stream>>>(h->dX, h->dinv_ell, h->dZ, h->D, e0, e1);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_kmat(b200bo_handle_s* h, double* dK, int64_t ld, int64_t N, int64_t Np, double noise, bool pad_identity) {
  const int64_t lim = pad_identity ? Np : N;
  const int T = (int)((lim + KT - 1) / KT);
  const int ntiles = T * (T + 1) / 2;
  if (ntiles == 0) return cudaSuccess;
  const size_t smem = (size_t)(2 * h->D * KT) * sizeof(double);
  const double sf2 = exp(2.0 * h->hp.lsigma);
#define B200BO_KMAT(F)                                                                                       \
  do {                                                                                                       \
    cudaFuncSetAttribute(kmat_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    kmat_kernel<F><<<ntiles, 256, smem, h->stream>>>(h->dZ, (int)N, (int)Np, h->D, sf2, noise, pad_identity ? 1 : 0, dK, ld); \
  } while (0)
  switch (h->fam) {
    case FAM_SE: B200BO_KMAT(FAM_SE); break;
    case FAM_MAT12: B200BO_KMAT(FAM_MAT12); break;
    case FAM_MAT32: B200BO_KMAT(FAM_MAT32); break;
    default: B200BO_KMAT(FAM_MAT52); break;
  }
#undef B200BO_KMAT
  h->launches++;
  return cudaGetLastError();
}

}  // namespace b200bo
