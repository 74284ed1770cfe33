// chol.cu -- K2-K4: blocked right-looking Cholesky of Sigma (FP64), in place on the lower triangle of h->dL.
//
// Replaces EXT `cholesky!(Symmetric(.., :U))` inside GaussianProcesses.jl `make_posdef!` (reached from the
// reference's update!(model, x, y), src/models/gp.jl:11-18).  The reference factor is the upper U with
// Sigma = U'U in column-major storage; that is byte-identical to the row-major lower L = U' kept here.
//
//   for each 128-wide panel k:
//     K2 potrf_diag   : L_kk = chol(A_kk) by one warp-cooperative CTA; the same sweep also produces L_kk^-1
//                       (forward substitution on I carried along the column loop) for the GEMM-shaped solves.
//     K3 trsm_panel   : A_ik <- A_ik L_kk^-T  = A_ik (L_kk^-1)^T          (DMMA GEMM, K = 128)
//     K4 syrk_trailing: A_ij <- A_ij - L_ik L_jk^T  for i >= j > k        (DMMA GEMM: the dense contraction)
//   The upper triangle of h->dL receives L^T ("mirrored factor") so the backward solve L^T w = v reads the
//   same k-major rows as the forward one.
#include "common.cuh"
#include "handle.h"

namespace b200bo {

// ------------------------------------------------------------------------------------------------------------
// K2: diagonal block.  S (128 x 129 doubles in smem): lower triangle = working matrix, later B (forward
// substitution applied to I); strict upper triangle receives L^T as columns are finished.
// ------------------------------------------------------------------------------------------------------------
constexpr int PS = NB + 1;

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, int64_t ld, int kb, double* __restrict__ Linv,
                                                            double* __restrict__ LinvT, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* S = sm;             // [NB][PS]
  double* dg = sm + NB * PS;  // [NB] diag(L)
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  double* Ab = A + ((int64_t)kb * NB) * ld + (int64_t)kb * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e >> 7, c = e & 127;
    if (c <= i) S[i * PS + c] = Ab[(int64_t)i * ld + c];
  }
  for (int j = 0; j <= NB; ++j) {
    __syncthreads();
    if (j < NB) {
      // ---- phase 1 of column j: pivot, L^T row j into the upper triangle, trailing rank-1 update ----
      double d = S[j * PS + j];
      if (!(d > 0.0)) {
        if (tid == 0) atomicCAS(info, 0, kb * NB + j + 1);
        d = 1.0;
      }
      const double piv = sqrt(d);
      const double rinv = 1.0 / piv;
      if (tid == 0) dg[j] = piv;
      for (int i = j + 1 + tid; i < NB; i += 256) S[j * PS + i] = S[i * PS + j] * rinv;
      for (int i = j + 1 + ty; i < NB; i += 16) {
        const double li = S[i * PS + j] * rinv;
        for (int k = j + 1 + tx; k <= i; k += 16) {
          const double lk = S[k * PS + j] * rinv;
          S[i * PS + k] = fma(-li, lk, S[i * PS + k]);
        }
      }
    }
    if (j > 0) {
      // ---- phase 2 of column jj = j-1: B[i][c] -= l_{i,jj} * B[jj][c] / l_{jj,jj},  i > jj, c <= jj  (B diag = 1) ----
      const int jj = j - 1;
      const double inv = 1.0 / dg[jj];
      for (int i = jj + 1 + ty; i < NB; i += 16) {
        const double lij = S[jj * PS + i] * inv;      // l_{i,jj} / l_{jj,jj}
        for (int c = tx; c <= jj; c += 16) {
          if (c == jj) S[i * PS + c] = -lij;
          else S[i * PS + c] = fma(-lij, S[jj * PS + c], S[i * PS + c]);
        }
      }
    }
  }
  __syncthreads();
  // write back: mirrored factor block, L^-1 and its transpose
  double* Li = Linv + (int64_t)kb * NB * NB;
  double* LiT = LinvT + (int64_t)kb * NB * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e >> 7, c = e & 127;
    const int lo = i < c ? i : c, hi = i < c ? c : i;
    Ab[(int64_t)i * ld + c] = (i == c) ? dg[i] : S[lo * PS + hi];
    const double inv_i = 1.0 / dg[i];
    const double y = (c < i) ? S[i * PS + c] * inv_i : (c == i ? inv_i : 0.0);   // Linv[i][c]
    Li[i * NB + c] = y;
  }
  for (int e = tid; e < NB * NB; e += 256) {
    const int c = e >> 7, i = e & 127;      // LinvT[c][i] = Linv[i][c]
    const double inv_i = 1.0 / dg[i];
    const double y = (c < i) ? S[i * PS + c] * inv_i : (c == i ? inv_i : 0.0);
    LiT[c * NB + i] = y;
  }
}

// ------------------------------------------------------------------------------------------------------------
// K3: panel solve.  One CTA per 64 rows below the diagonal block:  C = A_rows,k * Linv_kk^T  (in place),
// plus the mirrored copy into the upper triangle.
// ------------------------------------------------------------------------------------------------------------
constexpr int TR_BM = 64, TR_BN = 128, TR_STAGES = 3, TR_THREADS = 256;

__global__ void __launch_bounds__(TR_THREADS, 1) trsm_panel_kernel(double* __restrict__ A, int64_t ld, int kb,
                                                                   const double* __restrict__ Linv) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;   // 2 x 4 warps, warp tile 32 x 32
  const int64_t row0 = (int64_t)(kb + 1) * NB + (int64_t)blockIdx.x * TR_BM;
  double* Arow = A + row0 * ld + (int64_t)kb * NB;
  const double* Li = Linv + (int64_t)kb * NB * NB;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  gemm_mainloop<TR_BM, TR_BN, 4, 4, TR_STAGES, TR_THREADS, false>(acc, Arow, ld, Li, NB, NB / KC, sm, nullptr, 0, wm, wn, lane, tid,
                                                                 0, NB / KC);
  const int g = lane >> 2, q = lane & 3;
  double* Aup = A + ((int64_t)kb * NB) * ld + row0;    // mirrored block: rows = panel cols, cols = these rows
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int m = wm * 32 + mt * 8 + g, n = wn * 32 + nt * 8 + 2 * q;
      *reinterpret_cast<double2*>(Arow + (int64_t)m * ld + n) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
      Aup[(int64_t)n * ld + m] = acc[mt][nt][0];
      Aup[(int64_t)(n + 1) * ld + m] = acc[mt][nt][1];
    }
}

// ------------------------------------------------------------------------------------------------------------
// K4: trailing update (the dense contraction).  One CTA per 128 x 128 tile (bi >= bj > kb) of the trailing
// lower triangle:  A_ij -= L_ik L_jk^T, K = 128.
// ------------------------------------------------------------------------------------------------------------
constexpr int SY_BM = 128, SY_BN = 128, SY_STAGES = 3, SY_THREADS = 512;

__global__ void __launch_bounds__(SY_THREADS, 1) syrk_trailing_kernel(double* __restrict__ A, int64_t ld, int kb) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;   // 4 x 4 warps, warp tile 32 x 32
  const int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int bi = kb + 1 + ti, bj = kb + 1 + tj;
  const double* Ai = A + ((int64_t)bi * NB) * ld + (int64_t)kb * NB;
  const double* Aj = A + ((int64_t)bj * NB) * ld + (int64_t)kb * NB;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  gemm_mainloop<SY_BM, SY_BN, 4, 4, SY_STAGES, SY_THREADS, false>(acc, Ai, ld, Aj, ld, NB / KC, sm, nullptr, 0, wm, wn, lane, tid, 0,
                                                                 NB / KC);
  const int g = lane >> 2, q = lane & 3;
  double* C = A + ((int64_t)bi * NB) * ld + (int64_t)bj * NB;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int m = wm * 32 + mt * 8 + g, n = wn * 32 + nt * 8 + 2 * q;
      if (bi != bj || n + 1 <= m) {
        double2* p = reinterpret_cast<double2*>(C + (int64_t)m * ld + n);
        double2 c = *p;
        c.x -= acc[mt][nt][0];
        c.y -= acc[mt][nt][1];
        *p = c;
      } else if (n == m) {
        C[(int64_t)m * ld + n] -= acc[mt][nt][0];
      }
    }
}

cudaError_t launch_cholesky(b200bo_handle_s* h) {
  const int nblk = (int)(h->Np / NB);
  const size_t sm_potrf = (size_t)(NB * PS + NB) * sizeof(double);
  const size_t sm_trsm = (size_t)TR_STAGES * (TR_BM + TR_BN) * KC * sizeof(double);
  const size_t sm_syrk = (size_t)SY_STAGES * (SY_BM + SY_BN) * KC * sizeof(double);
  cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_potrf);
  cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_trsm);
  cudaFuncSetAttribute(syrk_trailing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_syrk);
  cudaMemsetAsync(h->dinfo, 0, sizeof(int), h->stream);
  while ((int)h->syrk_ev.size() < 2 * nblk) { cudaEvent_t e; cudaEventCreate(&e); h->syrk_ev.push_back(e); }
  h->syrk_ev_used = 0;
  for (int k = 0; k < nblk; ++k) {
    potrf_diag_kernel<<<1, 256, sm_potrf, h->stream>>>(h->dL, h->ld, k, h->dLinv, h->dLinvT, h->dinfo);
    h->launches++;
    const int rem = nblk - k - 1;
    if (rem > 0) {
      trsm_panel_kernel<<<rem * (NB / TR_BM), TR_THREADS, sm_trsm, h->stream>>>(h->dL, h->ld, k, h->dLinv);
      cudaEventRecord(h->syrk_ev[h->syrk_ev_used++], h->stream);
      syrk_trailing_kernel<<<rem * (rem + 1) / 2, SY_THREADS, sm_syrk, h->stream>>>(h->dL, h->ld, k);
      cudaEventRecord(h->syrk_ev[h->syrk_ev_used++], h->stream);
      h->launches += 2;
    }
  }
  return cudaGetLastError();
}

}  // namespace b200bo
