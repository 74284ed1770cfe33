// append.cu -- elastic append on the device: the rank-1 step (forward solve per point) and the blocked rank-m step from W = L^-1.
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

// ---- blocked append of m <= 16 points inside one 128-block, from the explicit inverse factor -------------------------------------------
// The acquisition path keeps W = L^-1 (h->dKi, acq_i8.cu), so the m new factor rows are ONE dense product instead of m chained forward
// solves:  V = Ks W^T  (Ks[j] = k(X, x_new_j)),  then  L22 = chol(K(Xnew, Xnew) + noise I - V V^T)  (m x m, one thread),  z and the rows
// of the last block's inverse follow as in the rank-1 step.  U12 = U11^-T A12 / U22 = chol(A22 - U12'U12) of EXT ElasticPDMats append!.
constexpr int AP_M = 16;

template <int FAM>
__global__ void kstar_rows_kernel(const double* __restrict__ Z, int N, int Np, int D, int m, double sf2, double* __restrict__ Ks, int64_t ldk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= Np || j >= m) return;
  double v = 0.0;
  if (i < N) {
    double r2 = 0.0;
    for (int d = 0; d < D; ++d) { const double df = Z[(int64_t)i * D + d] - Z[(int64_t)(N + j) * D + d]; r2 = fma(df, df, r2); }
    v = sf2 * kern_phi<FAM>(r2);
  }
  Ks[(int64_t)j * ldk + i] = v;
}

// V[j][i] = sum_{k <= i} W[i][k] Ks[j][k]: one warp per row i of W, fixed summation order
__global__ void __launch_bounds__(256) append_rows_kernel(const double* __restrict__ W, int64_t ld, const double* __restrict__ Ks, int64_t ldk, int N,
                                                          int m, double* __restrict__ V, double* __restrict__ L) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= N) return;
  double acc[AP_M];
#pragma unroll
  for (int j = 0; j < AP_M; ++j) acc[j] = 0.0;
  const double* wr = W + (int64_t)i * ld;
  for (int k = lane; k <= i; k += 32) {
    const double w = wr[k];
#pragma unroll
    for (int j = 0; j < AP_M; ++j) if (j < m) acc[j] = fma(w, Ks[(int64_t)j * ldk + k], acc[j]);
  }
  double mine = 0.0;
#pragma unroll
  for (int j = 0; j < AP_M; ++j) {
    double a = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0 && j < m) V[(int64_t)j * ldk + i] = a;
    if (lane == j) mine = a;
  }
  if (lane < m) L[(int64_t)i * ld + N + lane] = mine;         // the mirror columns of the new rows: m consecutive entries of row i
}

// Gram matrix of the new rows and their products with z: one CTA per pair (j >= c) / per row, fixed summation order.
// out[p]: p < m -> V_p . z;  then the pairs (0,0), (1,0), (1,1), (2,0), ...
__global__ void __launch_bounds__(256) append_gram_kernel(const double* __restrict__ V, int64_t ldk, const double* __restrict__ z, int N, int m,
                                                          double* __restrict__ out) {
  __shared__ double red[256];
  const int p = blockIdx.x, tid = threadIdx.x;
  int j, c; const double* b;
  if (p < m) { j = p; c = -1; b = z; }
  else { int q = p - m; j = 0; while (q > j) { q -= j + 1; ++j; } c = q; b = V + (int64_t)c * ldk; }
  (void)c;
  const double* a = V + (int64_t)j * ldk;
  double s = 0.0;
  for (int i = tid; i < N; i += 256) s = fma(a[i], b[i], s);
  red[tid] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) out[p] = red[0];
}

template <int FAM>
__global__ void __launch_bounds__(256) append_block_finish_kernel(const double* __restrict__ gram, double* __restrict__ L, int64_t ld, int N, int m, const double* __restrict__ V, int64_t ldk,
                                                                  const double* __restrict__ Z, int D, double* __restrict__ z, const double* __restrict__ y,
                                                                  double beta, double sf2, double noise, double* __restrict__ Linv,
                                                                  double* __restrict__ LinvT, int* __restrict__ info) {
  __shared__ double G[AP_M][AP_M + 1], L22[AP_M][AP_M + 1], vz[AP_M], zn[AP_M];
  __shared__ double lrow[NB];
  const int tid = threadIdx.x;
  for (int p = tid; p < m * (m + 1) / 2 + m; p += 256) {
    if (p < m) vz[p] = gram[p];
    else { int q = p - m, j = 0; while (q > j) { q -= j + 1; ++j; } G[j][q] = gram[p]; }
  }
  __syncthreads();
  if (tid == 0) {          // m x m Cholesky of A22 - V V^T and the new entries of z = L^-1 (y - m)
    for (int j = 0; j < m; ++j) {
      for (int c = 0; c <= j; ++c) {
        double r2 = 0.0;
        for (int d = 0; d < D; ++d) { const double df = Z[(int64_t)(N + j) * D + d] - Z[(int64_t)(N + c) * D + d]; r2 = fma(df, df, r2); }
        double sacc = sf2 * kern_phi<FAM>(r2) + (c == j ? noise : 0.0) - G[j][c];
        for (int k = 0; k < c; ++k) sacc = fma(-L22[j][k], L22[c][k], sacc);
        if (c < j) L22[j][c] = sacc / L22[c][c];
        else {
          if (!(sacc > 0.0)) { atomicCAS(info, 0, N + j + 1); sacc = 1.0; }
          L22[j][j] = sqrt(sacc);
        }
      }
      double t = y[N + j] - beta - vz[j];
      for (int k = 0; k < j; ++k) t = fma(-L22[j][k], zn[k], t);
      zn[j] = t / L22[j][j];
      z[N + j] = zn[j];
    }
  }
  __syncthreads();
  // the m x m corner of the new rows (their first N entries are copied from V by the launcher, the mirror columns were written by
  // append_rows_kernel)
  for (int j = 0; j < m; ++j)
    if (tid <= j) { const double l = L22[j][tid]; L[(int64_t)(N + j) * ld + N + tid] = l; L[(int64_t)(N + tid) * ld + N + j] = l; }
  // rows r0 .. r0+m-1 of the last block's inverse, one after the other:  Linv[r][c] = -(1/l_rr) sum_{k=c}^{r-1} L[r][k] Linv[k][c]
  const int blk = N / NB, r0 = N % NB;
  double* Li = Linv + (int64_t)blk * NB * NB;
  double* LiT = LinvT + (int64_t)blk * NB * NB;
  for (int j = 0; j < m; ++j) {
    const int r = r0 + j;
    __syncthreads();
    if (tid < NB) lrow[tid] = tid < r0 ? V[(int64_t)j * ldk + blk * NB + tid] : (tid < r ? L22[j][tid - r0] : 0.0);
    __syncthreads();
    if (tid < NB) {
      const int c = tid;
      const double ip = 1.0 / L22[j][j];
      double val;
      if (c < r) {
        double s = 0.0;
        for (int k = c; k < r; ++k) s = fma(lrow[k], Li[k * NB + c], s);
        val = -s * ip;
      } else {
        val = (c == r) ? ip : 0.0;
      }
      Li[r * NB + c] = val;
      LiT[c * NB + r] = val;
    }
    __threadfence_block();
  }
}

// m <= 16 new points that all fall into the current 128-block; needs W = L^-1 of the CURRENT factor in h->dKi (h->wt_valid)
cudaError_t launch_append_block(b200bo_handle_s* h, double noise, int m) {
  const int N = (int)h->N, D = h->D;
  const double sf2 = exp(2.0 * h->hp.lsigma);
  const double beta = h->mean_kind == B200BO_MEAN_CONST ? h->hp.beta : 0.0;
  int Np = (int)h->Np;
  if (N == Np) {     // the new points open a fresh block
    pad_block_kernel<<<64, 256, 0, h->stream>>>(h->dL, h->ld, Np, h->dLinv, h->dLinvT, h->dalpha, h->dz);
    h->launches++;
    Np += NB;
  }
  const int64_t ldk = h->cap;
  double* Ks = h->dV;                              // [16][cap]
  double* V = h->dV + AP_M * ldk;                  // [16][cap]
  const dim3 gk((Np + 255) / 256, m);
#define B200BO_APB(F)                                                                                                                  \
  do {                                                                                                                                 \
    kstar_rows_kernel<F><<<gk, 256, 0, h->stream>>>(h->dZ, N, Np, D, m, sf2, Ks, ldk);                                                 \
    append_rows_kernel<<<(N + 7) / 8, 256, 0, h->stream>>>(h->dKi, h->ld, Ks, ldk, N, m, V, h->dL);                                    \
    cudaMemcpy2DAsync(h->dL + (int64_t)N * h->ld, sizeof(double) * h->ld, V, sizeof(double) * ldk, sizeof(double) * N, m,             \
                      cudaMemcpyDeviceToDevice, h->stream);                                                                            \
    append_gram_kernel<<<m * (m + 1) / 2 + m, 256, 0, h->stream>>>(V, ldk, h->dz, N, m, h->dpart);                                     \
    append_block_finish_kernel<F><<<1, 256, 0, h->stream>>>(h->dpart, h->dL, h->ld, N, m, V, ldk, h->dZ, D, h->dz, h->dy, beta, sf2, noise, h->dLinv, \
                                                           h->dLinvT, h->dinfo);                                                       \
  } while (0)
  switch (h->fam) {
    case FAM_SE: B200BO_APB(FAM_SE); break;
    case FAM_MAT12: B200BO_APB(FAM_MAT12); break;
    case FAM_MAT32: B200BO_APB(FAM_MAT32); break;
    default: B200BO_APB(FAM_MAT52); break;
  }
#undef B200BO_APB
  h->launches += 4;
  h->N = N + m;
  h->Np = Np;
  return cudaGetLastError();
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
