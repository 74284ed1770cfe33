// joint.cu -- the JOINT posterior sample of myrand(model, X::Matrix) (reference src/models/gp.jl:7 -> EXT GaussianProcesses.jl
// rand(gp, X): mu + chol(make_posdef!(Sigma_post)) eps with the full M x M posterior covariance Sigma_post = K** - V'V).
//
// Re-design: Sigma_post is never formed.  The M points are appended (noise-free) to the observations of a worker model and the
// AUGMENTED covariance [[Sigma, K*], [K*', K**]] is factorised by the ordinary fit kernels (K1 + the blocked tcgen05/DMMA Cholesky):
// the trailing M x M block of that factor IS chol(K** - V'V) (the Schur complement), so the sample is one triangular
// matrix-vector product with the Philox normals of the Thompson stream.  make_posdef!'s retry rule acts on the trailing block's
// diagonal only (jitter 1e-6 tr(Sigma_post)/M per failed attempt, at most 10).
#include "common.cuh"
#include "acqfn.cuh"
#include "handle.h"

namespace b200bo {

// diagonal of the trailing block: k(x, x) = sigma_f^2 exactly for every stationary family, plus the current jitter (no noise term)
__global__ void joint_diag_kernel(double* __restrict__ L, int64_t ld, int64_t n0, int64_t m, double value) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) L[(n0 + i) * ld + n0 + i] = value;
}

__global__ void joint_eps_kernel(double* __restrict__ eps, int64_t m, unsigned long long seed, int64_t idx_offset) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) eps[i] = philox_normal(seed, (unsigned long long)(idx_offset + i));
}

// out_i = mu_i + sum_{j <= i} L[n0 + i][n0 + j] eps_j: one warp per row, lane-strided partial sums added in a fixed tree order
__global__ void __launch_bounds__(256) joint_trmv_kernel(const double* __restrict__ L, int64_t ld, int64_t n0, int64_t m,
                                                         const double* __restrict__ mu, const double* __restrict__ eps,
                                                         double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= m) return;
  const double* row = L + (n0 + i) * ld + n0;
  double acc = 0.0;
  for (int64_t j = lane; j <= i; j += 32) acc = fma(row[j], eps[j], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[i] = mu[i] + acc;
}

cudaError_t launch_joint_diag(b200bo_handle_s* h, int64_t n0, int64_t m, double value) {
  if (m <= 0) return cudaSuccess;
  joint_diag_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(h->dL, h->ld, n0, m, value);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_joint_sample(b200bo_handle_s* h, int64_t n0, int64_t m, const double* dmu, double* deps, double* dout,
                                unsigned long long seed, int64_t idx_offset) {
  if (m <= 0) return cudaSuccess;
  joint_eps_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(deps, m, seed, idx_offset);
  joint_trmv_kernel<<<(unsigned)((m + 7) / 8), 256, 0, h->stream>>>(h->dL, h->ld, n0, m, dmu, deps, dout);
  h->launches += 2;
  return cudaGetLastError();
}

}  // namespace b200bo
