// append.cu -- elastic rank-1 append on the device.
//
// Replaces EXT ElasticPDMats.jl `append!` reached from update!(model::GPE{<:ElasticArray}, x, y) (reference
// src/models/gp.jl:11; called every BO iteration at src/BayesianOptimization.jl:194): U12 = U11^-T A12, U22 = chol(A22 - U12'U12),
// alpha and mll refreshed -- O(N^2) instead of the O(N^3) refactor.  In the row-major lower layout kept here the new point adds
// ONE row:  L[N][0:N] = v = L^-1 k(X, x_new),  L[N][N] = sqrt(sf2 + noise - v'v);  z[N] = (y_N - m - v'z)/L[N][N];  the inverse of
// the last diagonal block gains one row;  alpha = L^-T z.
#include "common.cuh"
#include "handle.h"

namespace b200bo {

template <int FAM>
__global__ void kstar_vec_kernel(const double* __restrict__ Z, int N, int Np, int D, double sf2, double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np) return;
  double v = 0.0;
  if (i < N) {
    double r2 = 0.0;
    for (int d = 0; d < D; ++d) { const double df = Z[(int64_t)i * D + d] - Z[(int64_t)N * D + d]; r2 = fma(df, df, r2); }
    v = sf2 * kern_phi<FAM>(r2);
  }
  w[i] = v;
}

// a fresh 128-row block of the padded factor: identity rows, zero mirror columns, identity inverse blocks, zero alpha / z
__global__ void pad_block_kernel(double* __restrict__ L, int64_t ld, int Np_old, double* __restrict__ Linv, double* __restrict__ LinvT,
                                 double* __restrict__ alpha, double* __restrict__ z) {
  const int Np_new = Np_old + NB;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)NB * Np_new; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / Np_new), c = (int)(e % Np_new);
    L[(int64_t)(Np_old + r) * ld + c] = (c == Np_old + r) ? 1.0 : 0.0;       // new rows
    if (c < Np_old) L[(int64_t)c * ld + Np_old + r] = 0.0;                     // mirror columns
  }
  const int b = Np_old / NB;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < NB * NB; e += gridDim.x * blockDim.x) {
    const double v = ((e >> 7) == (e & 127)) ? 1.0 : 0.0;
    Linv[(int64_t)b * NB * NB + e] = v;
    LinvT[(int64_t)b * NB * NB + e] = v;
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < NB; e += gridDim.x * blockDim.x) { alpha[Np_old + e] = 0.0; z[Np_old + e] = 0.0; }
}

// one CTA: the new pivot, the new row of L (and its mirror column), z[N], and the new row of the last block's inverse
__global__ void __launch_bounds__(256) append_finish_kernel(double* __restrict__ L, int64_t ld, int N, const double* __restrict__ v,
                                                            double* __restrict__ z, const double* __restrict__ y, double beta, double diag,
                                                            double* __restrict__ Linv, double* __restrict__ LinvT, int* __restrict__ info) {
  __shared__ double s0[256], s1[256];
  __shared__ double vl[NB];
  __shared__ double piv;
  const int tid = threadIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = tid; i < N; i += 256) { const double vi = v[i]; a = fma(vi, vi, a); b = fma(vi, z[i], b); }
  s0[tid] = a; s1[tid] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s0[tid] += s0[tid + o]; s1[tid] += s1[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    double d = diag - s0[0];
    if (!(d > 0.0)) { atomicCAS(info, 0, N + 1); d = 1.0; }
    piv = sqrt(d);
    L[(int64_t)N * ld + N] = piv;
    z[N] = (y[N] - beta - s1[0]) / piv;
  }
  const int blk = N / NB, r = N % NB;
  if (tid < NB) vl[tid] = tid < r ? v[blk * NB + tid] : 0.0;
  __syncthreads();
  for (int i = tid; i < N; i += 256) { const double vi = v[i]; L[(int64_t)N * ld + i] = vi; L[(int64_t)i * ld + N] = vi; }
  // row r of the block inverse:  Linv[r][c] = -(1/piv) sum_{k=c}^{r-1} L[N][blk*128+k] Linv[k][c],  Linv[r][r] = 1/piv
  if (tid < NB) {
    const int c = tid;
    double* Li = Linv + (int64_t)blk * NB * NB;
    double* LiT = LinvT + (int64_t)blk * NB * NB;
    const double ip = 1.0 / piv;
    double val;
    if (c < r) {
      double s = 0.0;
      for (int k = c; k < r; ++k) s = fma(vl[k], Li[k * NB + c], s);
      val = -s * ip;
    } else {
      val = (c == r) ? ip : 0.0;
    }
    Li[r * NB + c] = val;
    LiT[c * NB + r] = val;
  }
}

cudaError_t launch_append_one(b200bo_handle_s* h, double noise, bool last) {
  // preconditions (capi.cu): x_new / y_new already at row N of dX / dZ / dy; h->N is still the OLD count, factor valid
  const int N = (int)h->N, D = h->D;
  const double sf2 = exp(2.0 * h->hp.lsigma);
  const double beta = h->mean_kind == B200BO_MEAN_CONST ? h->hp.beta : 0.0;
  int Np = (int)h->Np;
  if (N == Np) {     // the new point opens a fresh block
    pad_block_kernel<<<64, 256, 0, h->stream>>>(h->dL, h->ld, Np, h->dLinv, h->dLinvT, h->dalpha, h->dz);
    h->launches++;
    Np += NB;
  }
  const int nblk_old = (N + NB - 1) / NB;          // blocks that hold old points
  double* v = h->dV;                               // scratch for the new factor row
#define B200BO_KSTAR(F) kstar_vec_kernel<F><<<(Np + 255) / 256, 256, 0, h->stream>>>(h->dZ, N, Np, D, sf2, h->dw)
  switch (h->fam) {
    case FAM_SE: B200BO_KSTAR(FAM_SE); break;
    case FAM_MAT12: B200BO_KSTAR(FAM_MAT12); break;
    case FAM_MAT32: B200BO_KSTAR(FAM_MAT32); break;
    default: B200BO_KSTAR(FAM_MAT52); break;
  }
#undef B200BO_KSTAR
  h->launches++;
  cudaError_t e = launch_forward_solve(h, h->dw, v, nblk_old);
  if (e != cudaSuccess) return e;
  append_finish_kernel<<<1, 256, 0, h->stream>>>(h->dL, h->ld, N, v, h->dz, h->dy, beta, sf2 + noise, h->dLinv, h->dLinvT, h->dinfo);
  h->launches++;
  h->N = N + 1;
  h->Np = Np;
  if (!last) return cudaGetLastError();            // alpha and the log-determinant are refreshed once, behind the last point of a batch
  e = launch_backward_solve(h, h->dz, h->dw, h->dalpha, Np / NB);
  if (e != cudaSuccess) return e;
  return launch_logdet_dot(h);
}

}  // namespace b200bo
