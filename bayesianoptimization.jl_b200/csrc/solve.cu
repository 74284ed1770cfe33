// solve.cu -- K5: alpha = Sigma^-1 (y - m) by two single-RHS triangular solves, logdet, r'alpha.
//
// Replaces EXT GaussianProcesses.jl `update_mll!`:  alpha = cK \ (y - mean);  mll = -(y'alpha + logdet + N log 2pi)/2
// (reached from update!(model,x,y), reference src/models/gp.jl:11-18; SURVEY App. A "Fit").
// HBM-bound: each solve reads the triangle once (4 N^2 bytes).  Right-looking over 128-wide blocks, one launch
// per block: CTA b applies the just-solved block to its 128 rows; the CTA that owns the next diagonal block
// then solves it with the pre-inverted block (fixed summation order => deterministic).
#include <cstdlib>
#include "common.cuh"
#include "handle.h"

namespace b200bo {

// asm volatile loads keep their program order, so a batch really is in flight before its first use (the scheduler otherwise
// re-serialises load/use pairs to save registers, turning one memory latency into sixteen)
__device__ __forceinline__ double2 ldcg2_issue(const double* p) {
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ldcg1_issue(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];\n" : "=d"(v) : "l"(p));
  return v;
}

__global__ void residual_kernel(const double* __restrict__ y, double beta, double* __restrict__ w, int N, int Np) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np) w[i] = i < N ? y[i] - beta : 0.0;
}

// z_blk = M * w_blk with M given TRANSPOSED (MT[c][m]) so that threads m read coalesced.  256 threads: two halves of the c
// range, fixed combine order (deterministic).  The operand block is PREFETCHED into registers (diag_prefetch) before the
// caller's update phase, so the solve itself touches no global memory except w.
__device__ __forceinline__ void diag_prefetch(const double* __restrict__ MT, double (&v)[64], int tid) {
  const int m = tid & (NB - 1), half = tid >> 7;
#pragma unroll
  for (int u = 0; u < 64; ++u) v[u] = ldcg1_issue(MT + (half * 64 + u) * NB + m);
}
__device__ __forceinline__ void diag_apply(const double (&v)[64], const double* wblk, double* out, double* sh, int tid) {
  __shared__ double dpart[2][NB];
  if (tid < NB) sh[tid] = wblk[tid];
  __syncthreads();
  const int m = tid & (NB - 1), half = tid >> 7;
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < 64; ++u) s = fma(v[u], sh[half * 64 + u], s);
  dpart[half][m] = s;
  __syncthreads();
  if (tid < NB) out[tid] = dpart[0][tid] + dpart[1][tid];
}

// forward step i (i = -1: only the first diagonal solve).  grid = nblk - i - 1 CTAs (min 1), 256 threads.
__global__ void __launch_bounds__(256) fwd_step_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ LinvT,
                                                       double* __restrict__ w, double* __restrict__ z, int i) {
  __shared__ double sh[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rb = i + 1 + blockIdx.x;     // row block handled by this CTA
  double dv[64];
  if (blockIdx.x == 0) diag_prefetch(LinvT + (int64_t)rb * NB * NB, dv, tid);
  if (i >= 0) {
    // w[r] -= L[r][i*NB .. +127] . z_i ; one warp per row, 16 rows per warp, all 32 loads of a lane in flight at once
    double2 a[16], b[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const double* Lr = L + ((int64_t)rb * NB + warp + 8 * u) * ld + (int64_t)i * NB;
      a[u] = ldcg2_issue(Lr + 4 * lane);
      b[u] = ldcg2_issue(Lr + 4 * lane + 2);
    }
    if (tid < NB) sh[tid] = z[i * NB + tid];
    __syncthreads();
    const double z0 = sh[4 * lane], z1 = sh[4 * lane + 1], z2 = sh[4 * lane + 2], z3 = sh[4 * lane + 3];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      double s = a[u].x * z0;
      s = fma(a[u].y, z1, s);
      s = fma(b[u].x, z2, s);
      s = fma(b[u].y, z3, s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) w[(int64_t)rb * NB + warp + 8 * u] -= s;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0) diag_apply(dv, w + (int64_t)rb * NB, z + (int64_t)rb * NB, sh, tid);
}

// backward step i (descending; i = nblk: only the last diagonal solve).  grid = max(i,1) CTAs.
//   w[c] -= sum_m L[i*NB+m][c] * alpha_i[m] for the 128 columns c of column block cb; CTA cb == i-1 then solves it.
__global__ void __launch_bounds__(256) bwd_step_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ Linv,
                                                       double* __restrict__ w, double* __restrict__ alpha, int i, int nblk) {
  __shared__ double sh[NB];
  __shared__ double part[2][NB];
  const int tid = threadIdx.x;
  const int cb = (i < nblk) ? (int)blockIdx.x : nblk - 1;
  const bool solver = (i == nblk || cb == i - 1);      // alpha_cb = L_cb^-T w_cb : (Linv^T) given transposed == Linv
  double dv[64];
  if (solver) diag_prefetch(Linv + (int64_t)cb * NB * NB, dv, tid);
  if (i < nblk) {
    const int c = tid & 127, half = tid >> 7;   // two halves of the 128 rows, fixed combine order
    const double* Lc = L + ((int64_t)i * NB + half * 64) * ld + (int64_t)cb * NB + c;
    double v[64];
#pragma unroll
    for (int u = 0; u < 64; ++u) v[u] = ldcg1_issue(Lc + (int64_t)u * ld);       // 64 independent loads in flight
    if (tid < NB) sh[tid] = alpha[i * NB + tid];
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 64; ++u) s = fma(v[u], sh[half * 64 + u], s);
    part[half][c] = s;
    __syncthreads();
    if (tid < NB) w[(int64_t)cb * NB + tid] -= part[0][tid] + part[1][tid];
    __syncthreads();
  }
  if (solver) diag_apply(dv, w + (int64_t)cb * NB, alpha + (int64_t)cb * NB, sh, tid);
}

// The whole backward solve alpha = L^-T z in ONE launch: CTA b owns column block cb = nblk - 1 - b and stays resident; it streams
// the blocks L[i][cb] (i = nblk-1 .. cb+1) into registers one block ahead, waits for alpha_i through a release/acquire flag in
// global memory, accumulates, and finally applies the pre-inverted diagonal block and publishes alpha_cb.  The dependency chain
// costs one flag hop + a 128 x 128 product per block instead of one kernel launch (bwd_step_kernel: 32 dependent launches at
// N = 4096).  Deadlock-free: a CTA only waits on CTAs with a SMALLER blockIdx (dispatched no later than itself), nblk <= 148 CTAs
// of 512 threads are co-resident, and the spin is bounded (trap).  512 threads: thread (c, qd) covers rows 32 qd .. +31 of every
// block for column c; fixed summation order => deterministic.
// FWD = true is the forward solve z = L^-1 w with the same code: CTA b owns row block b, walks i = 0 .. b-1 and reads the blocks
// L[b][i] through the MIRRORED upper triangle (element (i*128 + r, b*128 + c) = L[b*128 + c][i*128 + r]), so the address pattern,
// the coalescing and the in-thread reduction are identical; `Minv` is LinvT there.
template <bool FWD>
__global__ void __launch_bounds__(512, 1) tri_persistent_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ Linv,
                                                                const double* __restrict__ z, double* __restrict__ alpha, int nblk,
                                                                int* __restrict__ flags, int epoch) {
  __shared__ double sh[NB];
  __shared__ double part[4][NB];
  const int tid = threadIdx.x, c = tid & 127, qd = tid >> 7;
  const int cb = FWD ? (int)blockIdx.x : nblk - 1 - (int)blockIdx.x;
  double dv[32], v[32];
  {
    const double* MT = Linv + (int64_t)cb * NB * NB;
#pragma unroll
    for (int u = 0; u < 32; ++u) dv[u] = ldcg1_issue(MT + (qd * 32 + u) * NB + c);
  }
  double acc = 0.0;
  for (int it = 0; it < (FWD ? cb : nblk - 1 - cb); ++it) {
    const int i = FWD ? it : nblk - 1 - it;
    const double* Lc = L + ((int64_t)i * NB + qd * 32) * ld + (int64_t)cb * NB + c;
#pragma unroll
    for (int u = 0; u < 32; ++u) v[u] = ldcg1_issue(Lc + (int64_t)u * ld);       // in flight while the flag is awaited
    if (tid == 0) {
      int f = 0, spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(f) : "l"(flags + i) : "memory");
      } while (f != epoch && ++spins < (1 << 26));
      if (f != epoch) __trap();
    }
    __syncthreads();
    if (tid < NB) sh[tid] = __ldcg(alpha + (int64_t)i * NB + tid);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 32; ++u) acc = fma(v[u], sh[qd * 32 + u], acc);
    __syncthreads();                                           // sh is rewritten in the next round
  }
  part[qd][c] = acc;
  __syncthreads();
  if (tid < NB) sh[tid] = z[(int64_t)cb * NB + tid] - (((part[0][tid] + part[1][tid]) + part[2][tid]) + part[3][tid]);
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < 32; ++u) s = fma(dv[u], sh[qd * 32 + u], s);
  part[qd][c] = s;
  __syncthreads();
  if (tid < NB) alpha[(int64_t)cb * NB + tid] = ((part[0][tid] + part[1][tid]) + part[2][tid]) + part[3][tid];
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(flags + cb), "r"(epoch) : "memory");
  }
}

// scal[0] = logdet = 2 sum log L_ii (i < N), scal[1] = r'alpha.  Single CTA, fixed order.
__global__ void __launch_bounds__(256) logdet_dot_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ y,
                                                         double beta, const double* __restrict__ alpha, int N, double* __restrict__ scal) {
  __shared__ double s0[256], s1[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) {
    a += log(L[(int64_t)i * ld + i]);
    b = fma(y[i] - beta, alpha[i], b);
  }
  s0[threadIdx.x] = a; s1[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { scal[0] = 2.0 * s0[0]; scal[1] = s1[0]; }
}

cudaError_t launch_residual(b200bo_handle_s* h, cudaStream_t st) {
  const double beta = h->mean_kind == B200BO_MEAN_CONST ? h->hp.beta : 0.0;
  residual_kernel<<<(int)((h->Np + 255) / 256), 256, 0, st>>>(h->dy, beta, h->dw, (int)h->N, (int)h->Np);
  h->launches++;
  return cudaGetLastError();
}

// Step i of z = L^-1 dw.  It needs column block i of L (K3 of panel i) and the inverse of diagonal block i+1 (K2 of panel i+1),
// so the factorisation launches it on its second stream right after K2 of panel i+1: the forward solve costs no time of its own.
cudaError_t launch_fwd_step(b200bo_handle_s* h, cudaStream_t st, int i, int nblk) {
  const int grid = i < 0 ? 1 : nblk - i - 1;
  fwd_step_kernel<<<grid, 256, 0, st>>>(h->dL, h->ld, h->dLinvT, h->dw, h->dz, i);
  h->launches++;
  return cudaGetLastError();
}

// z = L^-1 w  (w is overwritten), nblk blocks of the current factor
cudaError_t launch_forward_solve(b200bo_handle_s* h, double* w, double* z, int nblk) {
  static const bool stepwise = getenv("B200BO_BWD_STEPWISE") != nullptr;       // developer knob: the one-launch-per-block path
  if (!stepwise && nblk >= 1 && nblk <= h->num_sms) {
    tri_persistent_kernel<true><<<nblk, 512, 0, h->stream>>>(h->dL, h->ld, h->dLinvT, w, z, nblk, h->dflags, ++h->solve_epoch);
    h->launches++;
    return cudaGetLastError();
  }
  for (int i = -1; i < nblk - 1; ++i) {
    const int grid = i < 0 ? 1 : nblk - i - 1;
    fwd_step_kernel<<<grid, 256, 0, h->stream>>>(h->dL, h->ld, h->dLinvT, w, z, i);
    h->launches++;
  }
  return cudaGetLastError();
}

// alpha = L^-T z  (w is scratch)
cudaError_t launch_backward_solve(b200bo_handle_s* h, const double* z, double* w, double* alpha, int nblk) {
  static const bool stepwise = getenv("B200BO_BWD_STEPWISE") != nullptr;       // developer knob: the one-launch-per-block path
  if (!stepwise && nblk >= 1 && nblk <= h->num_sms) {
    tri_persistent_kernel<false><<<nblk, 512, 0, h->stream>>>(h->dL, h->ld, h->dLinv, z, alpha, nblk, h->dflags, ++h->solve_epoch);
    h->launches++;
    return cudaGetLastError();
  }
  cudaMemcpyAsync(w, z, sizeof(double) * nblk * NB, cudaMemcpyDeviceToDevice, h->stream);
  for (int i = nblk; i >= 1; --i) {
    const int grid = (i < nblk) ? i : 1;
    bwd_step_kernel<<<grid, 256, 0, h->stream>>>(h->dL, h->ld, h->dLinv, w, alpha, i, nblk);
    h->launches++;
  }
  return cudaGetLastError();
}

cudaError_t launch_logdet_dot(b200bo_handle_s* h) {
  const double beta = h->mean_kind == B200BO_MEAN_CONST ? h->hp.beta : 0.0;
  logdet_dot_kernel<<<1, 256, 0, h->stream>>>(h->dL, h->ld, h->dy, beta, h->dalpha, (int)h->N, h->dscal);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_alpha_mll(b200bo_handle_s* h, bool have_z) {
  const int nblk = (int)(h->Np / NB);
  cudaError_t e = cudaSuccess;
  if (!have_z) {
    e = launch_residual(h, h->stream);
    if (e != cudaSuccess) return e;
    e = launch_forward_solve(h, h->dw, h->dz, nblk);                    // z = L^-1 (y - m) is kept for elastic appends
    if (e != cudaSuccess) return e;
  }
  e = launch_backward_solve(h, h->dz, h->dw, h->dalpha, nblk);
  if (e != cudaSuccess) return e;
  return launch_logdet_dot(h);
}

}  // namespace b200bo
