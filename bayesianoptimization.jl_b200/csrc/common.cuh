// common.cuh -- shared device helpers for libb200bo (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace b200bo {

constexpr int NB = 128;        // block size of the blocked factor (rows/cols per panel)
constexpr int KC = 16;         // k-chunk in doubles per pipeline stage (128 B per row)
constexpr int TILE_N = 64;     // candidate columns per acquisition tile

// ---- kernel families: k = sf2 * phi(r2), psi = -2 dphi/dr2 (oracle/gp_oracle.py:_phi_psi) ----
enum { FAM_SE = 0, FAM_MAT12 = 1, FAM_MAT32 = 2, FAM_MAT52 = 3 };

template <int FAM>
__device__ __forceinline__ double kern_phi(double r2) {
  if (FAM == FAM_SE) return exp(-0.5 * r2);
  const double r = sqrt(r2);
  if (FAM == FAM_MAT12) return exp(-r);
  if (FAM == FAM_MAT32) { const double s = 1.7320508075688772 * r; return (1.0 + s) * exp(-s); }
  const double s = 2.23606797749979 * r;
  return (1.0 + s + s * s * (1.0 / 3.0)) * exp(-s);
}

template <int FAM>
__device__ __forceinline__ void kern_phi_psi(double r2, double& phi, double& psi) {
  if (FAM == FAM_SE) { phi = exp(-0.5 * r2); psi = phi; return; }
  const double r = sqrt(r2);
  if (FAM == FAM_MAT12) { phi = exp(-r); psi = r > 0.0 ? phi / r : 0.0; return; }
  if (FAM == FAM_MAT32) { const double s = 1.7320508075688772 * r; const double e = exp(-s); phi = (1.0 + s) * e; psi = 3.0 * e; return; }
  const double s = 2.23606797749979 * r; const double e = exp(-s);
  phi = (1.0 + s + s * s * (1.0 / 3.0)) * e; psi = (5.0 / 3.0) * (1.0 + s) * e;
}

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  lane = 4*g + q:
//      a = A[g][q], b = B[q][g], c0/c1 = C[g][2q], C[g][2q+1]  (SASS: DMMA.8x8x4) ----
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- cp.async (LDGSTS) 16-byte copies, L2-only (.cg) so same-CTA global writes are seen after bar.sync ----
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

// ---- shared-memory operand tiles are [row][k] with k contiguous, 16-byte chunks XOR-swizzled by the row so that
//      the LDS.128 fragment loads of a quarter-warp (rows r, r+1; 4 consecutive chunks) hit all 32 banks once ----
__device__ __forceinline__ int swz(int row) { return ((row & 1) << 2) | ((row >> 1) & 3); }
// offset (in doubles) of element (row, k) in a tile with `stride` doubles per row (stride % 16 == 0)
__device__ __forceinline__ int tile_off(int row, int k, int stride) {
  return row * stride + ((((k >> 1) ^ swz(row))) << 1) + (k & 1);
}

// stage one k-chunk (ROWS x KC doubles) of a k-major global matrix into a swizzled [ROWS][KC] tile
template <int ROWS, int NTHREADS>
__device__ __forceinline__ void stage_load(double* s, const double* __restrict__ g, int64_t ld, int tid) {
#pragma unroll
  for (int c = tid; c < ROWS * (KC / 2); c += NTHREADS) {
    const int r = c >> 3, ch = c & 7;
    cp_async16(s + r * KC + ((ch ^ swz(r)) << 1), g + (int64_t)r * ld + (ch << 1));
  }
}

// One k-chunk of warp-level MMA.  Warp tile = (MT*8) x (NT*8); a-rows start at arow0, b-rows at brow0.
// sA: swizzled tile with astride doubles per row, k offset ak0 (multiple of 8 in chunk units handled by caller)
template <int MT, int NT>
__device__ __forceinline__ void warp_mma_chunk(double (&acc)[MT][NT][2], const double* sA, int astride, int ak0, int arow0,
                                               const double* sB, int bstride, int bk0, int brow0, int lane) {
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int kk = 0; kk < KC / 8; ++kk) {
    double2 a[MT], b[NT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r = arow0 + mt * 8 + g;
      a[mt] = *reinterpret_cast<const double2*>(sA + tile_off(r, ak0 + kk * 8 + 2 * q, astride));
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int r = brow0 + nt * 8 + g;
      b[nt] = *reinterpret_cast<const double2*>(sB + tile_off(r, bk0 + kk * 8 + 2 * q, bstride));
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
        dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
      }
  }
}

// Multi-stage cp.async pipeline:  acc += sum_{c in [0,nchunks)} A[:, c*KC..] * B[:, c*KC..]^T
//   A: BM rows (k-major, ld = lda) from global; B: BN rows from global (ldb), or resident in smem (B_RES) as a
//   swizzled tile with bres_stride doubles per row starting at k = 0.
//   A warp (wm, wn) owns rows [wm*MT*8, ...) x cols [wn*NT*8, ...).  klo/khi (in chunks) bound the chunks this
//   warp needs (triangular operands); chunks outside are still staged but not multiplied.
template <int BM, int BN, int MT, int NT, int NSTAGE, int NTHREADS, bool B_RES>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[MT][NT][2], const double* __restrict__ gA, int64_t lda,
                                              const double* __restrict__ gB, int64_t ldb, int nchunks, double* stages,
                                              const double* sBres, int bres_stride, int wm, int wn, int lane, int tid,
                                              int klo, int khi) {
  constexpr int STAGE_DBL = (BM + (B_RES ? 0 : BN)) * KC;
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nchunks) {
      double* st = stages + s * STAGE_DBL;
      stage_load<BM, NTHREADS>(st, gA + s * KC, lda, tid);
      if (!B_RES) stage_load<BN, NTHREADS>(st + BM * KC, gB + s * KC, ldb, tid);
    }
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    const int pf = c + NSTAGE - 1;
    if (pf < nchunks) {
      double* st = stages + (pf % NSTAGE) * STAGE_DBL;
      stage_load<BM, NTHREADS>(st, gA + (int64_t)pf * KC, lda, tid);
      if (!B_RES) stage_load<BN, NTHREADS>(st + BM * KC, gB + (int64_t)pf * KC, ldb, tid);
    }
    cp_async_commit();
    if (c >= klo && c < khi) {
      const double* st = stages + (c % NSTAGE) * STAGE_DBL;
      if (B_RES)
        warp_mma_chunk<MT, NT>(acc, st, KC, 0, wm * MT * 8, sBres, bres_stride, c * KC, wn * NT * 8, lane);
      else
        warp_mma_chunk<MT, NT>(acc, st, KC, 0, wm * MT * 8, st + BM * KC, KC, 0, wn * NT * 8, lane);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

}  // namespace b200bo
