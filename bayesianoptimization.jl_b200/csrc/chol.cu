// chol.cu -- K2-K4: blocked right-looking Cholesky of Sigma (FP64), in place on the lower triangle of h->dL.
//
// Replaces EXT `cholesky!(Symmetric(.., :U))` inside GaussianProcesses.jl `make_posdef!` (reached from the
// reference's update!(model, x, y), src/models/gp.jl:11-18).  The reference factor is the upper U with
// Sigma = U'U in column-major storage; that is byte-identical to the row-major lower L = U' kept here.
//
//   for each 128-wide panel k:
//     K2 potrf_diag   : L_kk = chol(A_kk) and L_kk^-1 by ONE 256-thread CTA (latency-first; see the kernel's header);
//        chol_mirror  : the transposed triangles of that block (L^T mirror, row-major L^-1), off the chain.
//     K3 trsm_panel   : A_ik <- A_ik L_kk^-T = (L_kk^-1 A_ik^T)^T                (TMA + DMMA tile GEMM, K = 128)
//     K4 syrk_trailing: A_ij <- A_ij - L_ik L_jk^T  for i >= j > k              (inner updates: TMA + DMMA tile GEMM, K = 128)
//   The upper triangle of h->dL receives L^T ("mirrored factor") so the backward solve L^T w = v reads the same
//   k-major rows as the forward one.
//   Two-level blocking: inner panels of 128 inside outer panels of 512; the update right of a full outer panel is a K = 512
//   launch on the 5th-generation tensor cores (syrk_i8.cu: int8-slice product on tcgen05.mma, TMEM accumulators) -- the DMMA
//   kernel serves ragged panels and `b200bo_set_syrk_engine(h, 0)`.
//   Schedule (round 2, launch_cholesky_lookahead): only K2 and the HEAD of each panel (chol_head_kernel: the 128-row slab of K3 that
//   the next diagonal block needs + its rank-128 update, one 8-CTA cluster) sit on the critical chain; the rest of K3 / K4, the int8
//   slicing, the near and far K = 512 updates and the forward solve of y - m run on four side streams underneath the next K2, ordered by
//   events so every tile gets its updates in one fixed order.  The whole factorisation is captured into a CUDA graph per shape.  The
//   in-order schedule of round 1 stays selectable (knob "chol_sched" = 0) for A/B timing.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cooperative_groups.h>
#include "tma.cuh"
#include "tilegemm.cuh"
#include "handle.h"

namespace cg = cooperative_groups;

namespace b200bo {

// ------------------------------------------------------------------------------------------------------------
// K2: diagonal block (128 x 128) by ONE 256-thread CTA, latency-first.  The block lives in shared memory (row stride 132 doubles:
// DMMA fragment loads are conflict-free); it is factored in four 32-column sub-panels:
//   F(s)   warp 0 factors the 32 x 32 diagonal sub-block in REGISTERS (lane = row, columns exchanged by shuffles: no barrier, no
//          shared-memory round trip on the 128-column dependency chain), while warps 1-7 finish the far part U2(s-1) of the
//          previous rank-32 update (look-ahead);
//   T(s)   warps 1 .. 3-s: one thread per row below solves X L11^T = A21 by substitution, warp 5 inverts L11 (lane = column) -- both
//          TRAILING F(s) column by column (colbar[j]: column j of the unit factor and its pivot are published) instead of waiting
//          for the sub-block; the remaining warps do the look-ahead work U2(s-1) / block row s-1 of L^-1;
//   U1(s)  all warps: rank-32 update of the NEXT sub-panel's 32 columns on the FP64 tensor pipe (DMMA.8x8x4, K = 32).
// Then the off-diagonal blocks of L^-1 (W_ij = -W_ii sum_k L_ik W_kj) as 32^3 DMMA products, and a coalesced write-back of the
// mirrored factor block, L^-1 and L^-T.  Shared-memory image: lower triangle = L, strictly upper triangle = (L^-1)^T, rinv = 1/l_ii.
// ------------------------------------------------------------------------------------------------------------
constexpr int PLD = 132;                 // shared-memory row stride of the block (doubles)
constexpr int PD_THREADS = 256;
constexpr int PSUB = 32;                 // sub-panel width
constexpr size_t PD_SMEM = (size_t)(NB * PLD + PSUB * PSUB + 3 * PSUB * PSUB + 2 * NB + 2 * PSUB + 2 + PSUB) * sizeof(double);

#ifdef POTRF_PROF
__device__ unsigned long long g_head_prof[8];
__device__ unsigned long long g_potrf_prof[16];
#define PROF_T(i) do { if (threadIdx.x == 0) { const long long _t = clock64(); g_potrf_prof[i] += (unsigned long long)(_t - t_prev); t_prev = _t; } } while (0)
#else
#define PROF_T(i) do {} while (0)
#endif

// C(8x8 at rows r0, cols q0 of S) -= sum_{k < 32} S[r0 + .][k0 + k] * S[q0 + .][k0 + k]   (one warp; the k-steps are split into
// two independent accumulator chains: the DMMA dependent-issue latency, not its throughput, bounds a single tile)
__device__ __forceinline__ void rank32_tile(double* __restrict__ S, int r0, int q0, int k0, int g, int q) {
  double2* cp = reinterpret_cast<double2*>(S + (r0 + g) * PLD + q0 + 2 * q);
  double2 c = *cp;
  double e0 = 0.0, e1 = 0.0;
  const double* ap = S + (r0 + g) * PLD + k0 + q;
  const double* bp = S + (q0 + g) * PLD + k0 + q;
#pragma unroll
  for (int ks = 0; ks < PSUB / 4; ks += 2) {
    dmma884(c.x, c.y, -ap[4 * ks], bp[4 * ks]);
    dmma884(e0, e1, -ap[4 * ks + 4], bp[4 * ks + 4]);
  }
  c.x += e0; c.y += e1;
  *cp = c;
}


// two independent tiles at once (four accumulator chains in flight): the loads of both are issued before either product chain
__device__ __forceinline__ void rank32_tile_pair(double* __restrict__ S, int r0a, int q0a, int r0b, int q0b, int k0, int g, int q) {
  double2* cpa = reinterpret_cast<double2*>(S + (r0a + g) * PLD + q0a + 2 * q);
  double2* cpb = reinterpret_cast<double2*>(S + (r0b + g) * PLD + q0b + 2 * q);
  double2 ca = *cpa, cb2 = *cpb;
  double ea0 = 0.0, ea1 = 0.0, eb0 = 0.0, eb1 = 0.0;
  const double* apa = S + (r0a + g) * PLD + k0 + q;
  const double* bpa = S + (q0a + g) * PLD + k0 + q;
  const double* apb = S + (r0b + g) * PLD + k0 + q;
  const double* bpb = S + (q0b + g) * PLD + k0 + q;
  double aa[PSUB / 4], ba[PSUB / 4], ab[PSUB / 4], bb[PSUB / 4];
#pragma unroll
  for (int ks = 0; ks < PSUB / 4; ++ks) { aa[ks] = apa[4 * ks]; ba[ks] = bpa[4 * ks]; ab[ks] = apb[4 * ks]; bb[ks] = bpb[4 * ks]; }
#pragma unroll
  for (int ks = 0; ks < PSUB / 4; ks += 2) {
    dmma884(ca.x, ca.y, -aa[ks], ba[ks]);
    dmma884(cb2.x, cb2.y, -ab[ks], bb[ks]);
    dmma884(ea0, ea1, -aa[ks + 1], ba[ks + 1]);
    dmma884(eb0, eb1, -ab[ks + 1], bb[ks + 1]);
  }
  ca.x += ea0; ca.y += ea1; cb2.x += eb0; cb2.y += eb1;
  *cpa = ca; *cpb = cb2;
}

// 1/d and 1/sqrt(d) for d > 0 (normal range): MUFU seed + ONE cubically convergent step (3 dependent FP64 operations instead of
// the 8 of the library routines -- these sit on the 128-column dependency chain of the diagonal block)
__device__ __forceinline__ double fast_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = d * y;
  const double e = fma(-h, y, 1.0);                  // 1 - d y^2
  const double t = fma(e, 0.375, 0.5) * e;           // e/2 + 3 e^2 / 8
  return fma(y, t, y);
}

// Block row i (1..3) of W = L^-1:  W_ij = -W_ii sum_{k=j}^{i-1} L_ik W_kj for j < i, as 32^3 DMMA products by `nw` warps (this warp
// is worker `wid`); `named` selects the barrier between the two product phases (warps 1-7 only, or the whole CTA).
// W[r][c] (r >= c) is kept at S[c][r] (the diagonal of S holds w_rr = 1/l_rr); L[r][c] (r > c) at S[r][c].
// `phases`: 1 = the products P_j only (they need block rows < i of W, not W_ii), 2 = the final products only, 3 = both.
__device__ __forceinline__ void winv_block_row(double* __restrict__ S, double* __restrict__ Pb, int i, int wid, int nw, bool named, int g,
                                               int q, int phases = 3) {
  const int nthr = 32 * nw;
  // P_j = sum_k L_ik W_kj  (32 x 32, K = 32 (i - j)); two accumulator chains per 8 x 8 tile
  if (phases & 1)
  for (int t = wid; t < i * 16; t += nw) {
    const int j = t >> 4, tr = (t >> 2) & 3, tc = t & 3;
    double c0a = 0.0, c1a = 0.0, e0a = 0.0, e1a = 0.0;
    const double* lrow = S + (PSUB * i + 8 * tr + g) * PLD + q;     // L[r][kk + q]
    const int cw = PSUB * j + 8 * tc + g;                            // column of W_kj read by this lane as B[k = q][n = g]
    const double* wcol = S + cw * PLD + q;                           // W[kk + q][cw]
#pragma unroll
    for (int kk = PSUB * j; kk < PSUB * j + PSUB; kk += 8) {         // k = j: W_jj is lower triangular
      const double w0 = wcol[kk], w1 = wcol[kk + 4];
      dmma884(c0a, c1a, lrow[kk], (kk + q >= cw) ? w0 : 0.0);
      dmma884(e0a, e1a, lrow[kk + 4], (kk + 4 + q >= cw) ? w1 : 0.0);
    }
#pragma unroll 2
    for (int kk = PSUB * (j + 1); kk < PSUB * i; kk += 8) {          // k > j: full blocks
      dmma884(c0a, c1a, lrow[kk], wcol[kk]);
      dmma884(e0a, e1a, lrow[kk + 4], wcol[kk + 4]);
    }
    double* pp = Pb + (j * PSUB + 8 * tr + g) * PSUB + 8 * tc + 2 * q;
    *reinterpret_cast<double2*>(pp) = make_double2(c0a + e0a, c1a + e1a);
  }
  if (phases == 3) { if (named) asm volatile("bar.sync 1, %0;\n" ::"r"(nthr) : "memory"); else __syncthreads(); }
  // W_ij = -W_ii P_j, stored transposed into the upper triangle
  if (phases & 2)
  for (int t = wid; t < i * 16; t += nw) {
    const int j = t >> 4, tr = (t >> 2) & 3, tc = t & 3;
    double c0a = 0.0, c1a = 0.0, e0a = 0.0, e1a = 0.0;
    const int r = PSUB * i + 8 * tr + g;
    const double* pcol = Pb + (j * PSUB + q) * PSUB + 8 * tc + g;    // P_j[kk + q][8 tc + g]
#pragma unroll
    for (int kk = 0; kk < PSUB; kk += 8) {
      const int k0 = PSUB * i + kk + q;
      const double w0 = S[k0 * PLD + r], w1 = S[(k0 + 4) * PLD + r]; // W_ii[r][k] at S[k][r], lower triangular
      dmma884(c0a, c1a, (r >= k0) ? -w0 : 0.0, pcol[kk * PSUB]);
      dmma884(e0a, e1a, (r >= k0 + 4) ? -w1 : 0.0, pcol[(kk + 4) * PSUB]);
    }
    const int cc = PSUB * j + 8 * tc + 2 * q;
    S[cc * PLD + r] = c0a + e0a;
    S[(cc + 1) * PLD + r] = c1a + e1a;
  }
}

__global__ void __launch_bounds__(PD_THREADS, 1) potrf_diag_kernel(double* __restrict__ A, int64_t ld, int kb, double* __restrict__ Linv,
                                                                   double* __restrict__ LinvT, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;                          // [NB][PLD]
  double* Mt = S + NB * PLD;               // [32][32] current diagonal sub-block, scaled and transposed: Mt[c*32 + r] = l_rc / l_cc
  double* Pb = Mt + PSUB * PSUB;           // [3][32][32] products of the block inverse
  double* rinv = Pb + 3 * PSUB * PSUB;     // [NB] 1 / l_ii
  double* ldiag = rinv + NB;               // [NB] l_ii
  double* cb = ldiag + NB;                 // [32] pivots d_j of the current sub-block, published column by column (+ 32 spare)
  uint64_t* bar = reinterpret_cast<uint64_t*>(cb + 2 * PSUB);
  uint64_t* colbar = bar + 1;              // [32] "column j of the current sub-block's unit factor is in Mt" (phase = s & 1)
  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler
  double* Ab = A + ((int64_t)kb * NB) * ld + (int64_t)kb * NB;
#ifdef POTRF_PROF
  long long t_prev = clock64();
#endif
  // ---- load the lower triangle: one 1-D TMA bulk copy per row (the strictly upper part is never read before it is written) ----
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (tid < PSUB) mbar_init(&colbar[tid], 1);      // generic-proxy waiters only: the __syncthreads below publishes them
  __syncthreads();
  if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(8 * (NB / 2) * (NB / 2 + 1) * 2));   // sum_i 8 * ((i + 2) & ~1)
  __syncthreads();
  if (tid < NB) {
    const uint32_t bytes = (uint32_t)(((tid + 2) & ~1) * 8);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(S + tid * PLD)),
                 "l"(Ab + (int64_t)tid * ld), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
  }
  mbar_wait(bar, 0);
  PROF_T(0);
#pragma unroll 1
  for (int s = 0; s < NB / PSUB; ++s) {
    const int c0 = s * PSUB;
    // Roles while the diagonal sub-block is factored.  warp 0: F(s).  warps 1 .. 3-s: T(s), one thread per row below, TRAILING the
    // factorisation column by column (column j of the unit factor is published through colbar[j] the moment F(s) has it).  warp 5: the
    // inverse of the sub-block, trailing likewise.  The other 3 + s warps: the look-ahead work U2(s-1) and block row s-1 of L^-1.
    const int nT = NB / PSUB - 1 - s;                  // warps that own rows below
    const uint32_t cpar = (uint32_t)s & 1u;
    double y[PSUB];                                    // T(s): this thread's row;  inverse: this lane's column
    if (warp == 0) {
      // ---------------- F(s): 32 x 32 factorisation in registers, lane = row, square-root free on the chain ----------------
      // a[j] holds the UNSCALED column c_ij = l_ij sqrt(d_j); the update is a_ik -= c_ij (c_kj / d_j).  The only serial chain is
      // d_j -> 1/d_j -> d_{j+1} (the diagonal stays in the owning lane); the scaled column c_kj / d_j = l_kj / l_jj reaches the other
      // lanes -- and the trailing warps -- through Mt (one STS + broadcast LDS.128 instead of 31 shuffle pairs per column).
      double a[PSUB];
      const double* row = S + (c0 + lane) * PLD + c0;
      // only the lower triangle is read: the strictly upper part of the sub-block is where the trailing inverse warp writes W^T
#pragma unroll
      for (int j = 0; j < PSUB; j += 2) {
        double2 v = make_double2(0.0, 0.0);
        if (j + 1 <= lane) v = *reinterpret_cast<const double2*>(row + j); else if (j == lane) v.x = row[j];
        a[j] = v.x; a[j + 1] = v.y;
      }
      double dg = 0.0;                                 // this lane's diagonal element, fully updated once column lane-1 is done
#pragma unroll
      for (int j = 0; j < PSUB; ++j) if (j == lane) dg = a[j];
      // Branch-free straight-line code: the next pivot's broadcast and reciprocal are issued BEFORE the bulk of the current
      // column's update, so the in-order warp overlaps the d -> 1/d chain with the independent FMAs.
      double my_d = 1.0;
      int bad = 64;
      double d = __shfl_sync(0xffffffffu, dg, 0);
      bad = (d > 0.0) ? bad : 0;
      d = (d > 0.0) ? d : 1.0;
      double rc = fast_rcp(d);
#pragma unroll
      for (int j = 0; j < PSUB; ++j) {
        double* cbj = Mt + j * PSUB;
        cbj[lane] = a[j] * rc;                        // c_ij / d_j  (meaningful for lanes i > j)
        my_d = (lane == j) ? d : my_d;
        dg = fma(-(a[j] * a[j]), rc, dg);             // lanes i > j (a[j]^2 is ready before rc is)
        double dn = 1.0, rcn = 1.0;
        if (j + 1 < PSUB) {
          dn = __shfl_sync(0xffffffffu, dg, j + 1);   // lane j+1's diagonal is final now
          bad = (dn > 0.0) ? bad : min(bad, j + 1);
          dn = (dn > 0.0) ? dn : 1.0;
          rcn = fast_rcp(dn);
        }
        __syncwarp();
        if (lane == 0) { cb[j] = d; mbar_arrive(&colbar[j]); }   // column j and its pivot d_j are published (release)
#pragma unroll
        for (int k = (j + 1) & ~1; k < PSUB; k += 2) {
          const double2 m = *reinterpret_cast<const double2*>(cbj + k);
          if (k > j) a[k] = fma(-a[j], m.x, a[k]);    // meaningful for lanes i > k
          a[k + 1] = fma(-a[j], m.y, a[k + 1]);
        }
        d = dn; rc = rcn;
      }
      if (bad < 64 && lane == 0) atomicCAS(info, 0, kb * NB + c0 + bad + 1);
      const double my_rs = fast_rsqrt(my_d);
      rinv[c0 + lane] = my_rs;
      ldiag[c0 + lane] = my_d * my_rs;                // l_ii = sqrt(d_i)
      __syncwarp();
      double* wrow = S + (c0 + lane) * PLD + c0;
#pragma unroll
      for (int j = 0; j < PSUB; ++j) {
        const double l = a[j] * rinv[c0 + j];         // l_ij
        if (j < lane) wrow[j] = l;
      }
      wrow[lane] = my_rs;                             // the diagonal of S carries w_ii = 1 / l_ii
    } else if (warp <= nT) {
      // ---------------- T(s), trailing: y_c = x_c l_cc;  y_c2 -= y_c (l_c2,c / l_cc) as soon as column c is out ----------------
      const int r = c0 + PSUB + (tid - 32);
      double* row = S + r * PLD + c0;
#pragma unroll
      for (int j = 0; j < PSUB; j += 2) { const double2 v = *reinterpret_cast<const double2*>(row + j); y[j] = v.x; y[j + 1] = v.y; }
      double xe = 0.0;
#pragma unroll
      for (int c = 0; c < PSUB; ++c) {
        mbar_wait(&colbar[c], cpar);
        const double x = y[c] * fast_rsqrt(cb[c]);    // l_rc = y_c / l_cc: final, stored in pairs (the trailing warp has the time for the rsqrt)
        if (c & 1) *reinterpret_cast<double2*>(row + c - 1) = make_double2(xe, x); else xe = x;
#pragma unroll
        for (int c2 = c + 1; c2 < PSUB; ++c2) y[c2] = fma(-y[c], Mt[c * PSUB + c2], y[c2]);
      }
    } else if (warp == 5) {
      // ---------------- W = L11^-1, trailing, lane = column c:  v_k = w_k l_kk = (k == c) ? 1 : acc_k,  acc_i -= (l_ik / l_kk) v_k ------
#pragma unroll
      for (int i = 0; i < PSUB; ++i) y[i] = 0.0;
      double* wrow = S + (c0 + lane) * PLD + c0;      // (W^T)[c][i] = W[i][c] -> strictly upper part of row c
#pragma unroll
      for (int k = 0; k < PSUB; ++k) {
        mbar_wait(&colbar[k], cpar);
        const double vk = (k == lane) ? 1.0 : y[k];
        if (k > lane) wrow[k] = vk * fast_rsqrt(cb[k]);
#pragma unroll
        for (int i = k + 1; i < PSUB; ++i) y[i] = fma(-Mt[k * PSUB + i], vk, y[i]);
      }
    } else {
      // ---------------- look-ahead workers: the warps of {1..7} that are neither T(s) rows nor the inverse ----------------
      const int nw = NB / PSUB - 1 + s;                                   // 3 + s
      const int wid = (warp - (nT + 1)) - (warp > 5 ? 1 : 0);
      if (s > 0) {
        // U2(s-1): far part of the previous rank-32 update (columns >= c0 + 32)
        const int k0 = c0 - PSUB, t0 = (c0 + PSUB) / 8, n = NB / 8 - t0;     // tile rows/cols t0 .. 15
        for (int t = wid; t < n * (n + 1) / 2; t += nw) {
          int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
          while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
          while (ti * (ti + 1) / 2 > t) --ti;
          const int tj = t - ti * (ti + 1) / 2;
          rank32_tile(S, 8 * (t0 + ti), 8 * (t0 + tj), k0, g, q);
        }
      }
      // block row s-1 of L^-1 (needs the inverses of diagonal sub-blocks 0 .. s-1): hidden behind F(s)
      if (s >= 2) winv_block_row(S, Pb, s - 1, wid, nw, true, g, q);
      if (s == NB / PSUB - 1) {   // the products of the LAST block row only need rows < 3 of W: also hidden behind F(3)
        asm volatile("bar.sync 1, %0;\n" ::"r"(32 * nw) : "memory");       // row s-1 of W complete, Pb free again
        winv_block_row(S, Pb, s, wid, nw, true, g, q, 1);
      }
    }
#ifdef POTRF_PROF
    if (threadIdx.x == 0) g_potrf_prof[8 + s] += (unsigned long long)(clock64() - t_prev);      // warp 0's own F(s) time
#endif
    __syncthreads();
#ifdef POTRF_PROF
    if (threadIdx.x == 0) g_potrf_prof[12 + s] += (unsigned long long)(clock64() - t_prev);     // the whole phase A(s)
#endif
    PROF_T(1);
    PROF_T(2);
    // ---------------- U1(s): rank-32 update of the next sub-panel's columns ----------------
    if (s + 1 < NB / PSUB) {
      const int t0 = (c0 + PSUB) / 8, nr = NB / 8 - t0;                    // tile rows t0 .. 15, tile cols t0 .. t0 + 3
      const int ntile = 10 + 4 * (nr - 4);
      auto tile_of = [&](int t, int& ti, int& tj) {
        if (t < 10) { ti = (t >= 6) ? 3 : (t >= 3) ? 2 : (t >= 1) ? 1 : 0; tj = t - ti * (ti + 1) / 2; }
        else { ti = 4 + ((t - 10) >> 2); tj = (t - 10) & 3; }
      };
      for (int t = warp; t < ntile; t += 2 * (PD_THREADS / 32)) {           // two tiles per step: t and t + 8
        int ti, tj, ui, uj;
        tile_of(t, ti, tj);
        if (t + PD_THREADS / 32 < ntile) {
          tile_of(t + PD_THREADS / 32, ui, uj);
          rank32_tile_pair(S, 8 * (t0 + ti), 8 * (t0 + tj), 8 * (t0 + ui), 8 * (t0 + uj), c0, g, q);
        } else {
          rank32_tile(S, 8 * (t0 + ti), 8 * (t0 + tj), c0, g, q);
        }
      }
      __syncthreads();
    }
    PROF_T(3);
  }
  // ---------------- the last block row of L^-1: only its final products are left ----------------
  winv_block_row(S, Pb, NB / PSUB - 1, warp, PD_THREADS / 32, false, g, q, 2);   // W_3j = -W_33 P_j (P_j computed during F(3))
  __syncthreads();
  PROF_T(4);
  // ---------------- write-back of the two triangles that shared memory holds row-wise: L (lower, with its diagonal) into the factor and
  // L^-T (upper) into LinvT -- 16-byte pieces, a warp store = 512 contiguous bytes of one row.  Their transposes (the mirrored upper
  // triangle of the factor block, L^-1) are written by chol_mirror_kernel OFF the critical chain: one SM stores ~64 B/clk, and the four
  // triangles together (393 KB) were 10.7 k of this kernel's 80 k cycles. ----------------
  {
    double* LiT = LinvT + (int64_t)kb * NB * NB;
#pragma unroll 4
    for (int p = tid; p < NB * (NB / 2); p += PD_THREADS) {
      const int i = p >> 6, c = 2 * (p & 63);
      const double2 t = *reinterpret_cast<const double2*>(S + i * PLD + c);
      if (c <= i) {                                    // L[i][c .. c+1]; beyond the diagonal the mirror kernel writes
        double2 v = t;
        if (c == i) v.x = ldiag[i];
        if (c + 1 == i) v.y = ldiag[i];
        if (c + 1 <= i) *reinterpret_cast<double2*>(Ab + (int64_t)i * ld + c) = v; else Ab[(int64_t)i * ld + c] = v.x;
      }
      if (c + 1 >= i) {                                // (W^T)[i][c .. c+1] = W[c .. c+1][i]: row i of S, diagonal w_ii included
        if (c >= i) *reinterpret_cast<double2*>(LiT + i * NB + c) = t; else LiT[i * NB + c + 1] = t.y;
      }
    }
  }
#ifdef POTRF_PROF
  __syncthreads();
  PROF_T(5);
#endif
}


// ------------------------------------------------------------------------------------------------------------
// Head of panel k (look-ahead schedule): everything the NEXT diagonal block needs from panel k, as ONE launch of an 8-CTA cluster.
//   phase 1   X = L_{k+1,k} = A_{k+1,k} W_k^T   (W_k = L_kk^-1, lower triangular: only l <= j contributes)
//   phase 2   A_{k+1,k+1} -= X X^T              (lower tiles; + the scratch accumulator D of an outer-panel boundary)
// CTA r owns the 8-row tiles r and 15 - r of the block row (balanced triangle); W_k^T (from L^-T) and the CTA's rows of A arrive by 1-D
// TMA bulk copies; every CTA writes its 16 rows of X to the factor, the cluster barrier orders those writes, and phase 2 gathers the rows
// it multiplies with back from L2 into shared memory in one sweep (reading them through distributed shared memory measured ~15 B/clk).
// The cluster is what makes the two phases ONE launch: its barrier is the grid-wide synchronisation between them.  All products on the
// FP64 tensor pipe (DMMA.8x8x4), four independent accumulator chains per warp.
// The tile GEMM kernels need ~12 us per launch for this (one 128 x 64 x 128 tile per CTA is 8.4 us of DMMA on one SM): twice that sat
// between every two diagonal blocks.
// ------------------------------------------------------------------------------------------------------------
constexpr int HC = 8;
constexpr size_t HEAD_SMEM = (size_t)(NB * PLD + 2 * 16 * PLD) * sizeof(double) + 16;

__global__ void __cluster_dims__(HC, 1, 1) __launch_bounds__(256, 1) chol_head_kernel(double* __restrict__ A, int64_t ld, int kb,
                                                                                       const double* __restrict__ LinvT,
                                                                                       const double* __restrict__ Dacc) {
  extern __shared__ __align__(16) double sm[];
  double* W = sm;                        // [128][PLD] upper triangle of W_k^T: W[j][l] at W[l * PLD + j]  (later: the gathered rows of X)
  double* As = W + NB * PLD;             // [16][PLD] this CTA's rows of A_{k+1,k}
  double* Xs = As + 16 * PLD;            // [16][PLD] this CTA's rows of X
  uint64_t* bar = reinterpret_cast<uint64_t*>(Xs + 16 * PLD);
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int b = kb + 1;
  const int mrow0 = 8 * r, mrow1 = 8 * (15 - r);
  double* Ablk = A + ((int64_t)b * NB) * ld;            // block row b
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(8 * (NB / 2) * (NB / 2 + 1) * 2 + 16 * NB * 8));
  __syncthreads();
  if (tid < NB) {                                      // row l = tid of W^T: entries j >= l (from the even column below l)
    const int c0 = tid & ~1;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(W + tid * PLD + c0)),
                 "l"(LinvT + (int64_t)kb * NB * NB + (int64_t)tid * NB + c0), "r"((uint32_t)((NB - c0) * 8)), "r"(smem_u32(bar))
                 : "memory");
  } else if (tid < NB + 16) {
    const int rr = tid - NB, row = (rr < 8 ? mrow0 : mrow1) + (rr & 7);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(As + rr * PLD)),
                 "l"(Ablk + (int64_t)row * ld + (int64_t)kb * NB), "r"((uint32_t)(NB * 8)), "r"(smem_u32(bar))
                 : "memory");
  }
  // phase-2 accumulators start from the diagonal block (issued now: the loads fly during phase 1).  Tile t of the CTA: t <= r ->
  // (m-tile r, column tile t), else (m-tile 15 - r, column tile t - r - 1); warp w owns t = w, w + 8, w + 16.
  double2 c[3];
  int tm[3], tj[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int t = warp + 8 * i;
    tm[i] = -1; tj[i] = 0; c[i] = make_double2(0.0, 0.0);
    if (t < 17) {
      tm[i] = t <= r ? 0 : 1;
      tj[i] = t <= r ? t : t - r - 1;
      const int row = (tm[i] ? mrow1 : mrow0) + g, col = 8 * tj[i] + 2 * q;
      c[i] = *reinterpret_cast<const double2*>(Ablk + (int64_t)row * ld + (int64_t)b * NB + col);
      if (Dacc) { const double2 d = *reinterpret_cast<const double2*>(Dacc + row * NB + col); c[i].x += d.x; c[i].y += d.y; }
    }
  }
#ifdef POTRF_PROF
  long long hp_t = clock64();
#define HPROF(i) do { if (tid == 0 && r == 0) { const long long _t = clock64(); g_head_prof[i] += (unsigned long long)(_t - hp_t); hp_t = _t; } } while (0)
#else
#define HPROF(i) do {} while (0)
#endif
  mbar_wait(bar, 0);
  HPROF(0);
  // ---- phase 1: X(16 x 128) = As W^T; warp w computes column tiles w and 15 - w for both m-tiles ----
#pragma unroll
  for (int ni = 0; ni < 2; ++ni) {
    const int jt = ni ? 15 - warp : warp;
    const double* brow = W + q * PLD + 8 * jt + g;     // B[k = q][n = g] = W[8 jt + g][l0 + q] = (W^T)[l0 + q][8 jt + g]
    const double* a0 = As + g * PLD + q;
    const double* a1 = As + (8 + g) * PLD + q;
    double2 x0 = make_double2(0.0, 0.0), x1 = x0, y0 = x0, y1 = x0;
#pragma unroll 4
    for (int ks = 0; ks < 2 * jt; ks += 2) {          // l < 8 jt: full k-steps, two chains per m-tile
      const double b0 = brow[4 * ks * PLD], b1 = brow[(4 * ks + 4) * PLD];
      dmma884(x0.x, x0.y, a0[4 * ks], b0);
      dmma884(x1.x, x1.y, a1[4 * ks], b0);
      dmma884(y0.x, y0.y, a0[4 * ks + 4], b1);
      dmma884(y1.x, y1.y, a1[4 * ks + 4], b1);
    }
    {                                                 // the two k-steps that straddle the diagonal of W: l <= j only
      const double b0 = (q <= g) ? brow[8 * jt * PLD] : 0.0, b1 = (4 + q <= g) ? brow[(8 * jt + 4) * PLD] : 0.0;
      dmma884(x0.x, x0.y, a0[8 * jt], b0);
      dmma884(x1.x, x1.y, a1[8 * jt], b0);
      dmma884(y0.x, y0.y, a0[8 * jt + 4], b1);
      dmma884(y1.x, y1.y, a1[8 * jt + 4], b1);
    }
    x0.x += y0.x; x0.y += y0.y; x1.x += y1.x; x1.y += y1.y;
    const int col = 8 * jt + 2 * q;
    *reinterpret_cast<double2*>(Xs + g * PLD + col) = x0;
    *reinterpret_cast<double2*>(Xs + (8 + g) * PLD + col) = x1;
    // L_{k+1,k} to the factor: lower triangle and its mirror
    *reinterpret_cast<double2*>(Ablk + (int64_t)(mrow0 + g) * ld + (int64_t)kb * NB + col) = x0;
    *reinterpret_cast<double2*>(Ablk + (int64_t)(mrow1 + g) * ld + (int64_t)kb * NB + col) = x1;
    double* up = A + ((int64_t)kb * NB + col) * ld + (int64_t)b * NB;
    up[mrow0 + g] = x0.x; up[ld + mrow0 + g] = x0.y;
    up[mrow1 + g] = x1.x; up[ld + mrow1 + g] = x1.y;
  }
  HPROF(1);
  cluster.sync();                                     // every CTA's rows of X are in its shared memory; W is dead from here on
  HPROF(2);
  // ---- gather the rows of X this CTA multiplies with (rows 0 .. 8 (16 - r) - 1 of the slab) into the W region.  They come back from
  //      L2, where every CTA has just written its rows of L_{k+1,k} (cluster.sync orders those writes; ld.global.cg bypasses L1): pulling
  //      them through distributed shared memory instead measured 8.8 k cycles for the 128 KB of CTA 0 (~15 B/clk) ----
  {
    const int npiece = (16 - r) * 8 * (NB / 2);        // 16-byte pieces: 64 per row
    const double* Xg = Ablk + (int64_t)kb * NB;        // row i of the slab at Xg + i * ld
    for (int p = tid; p < npiece; p += 8 * 256) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int pp = p + u * 256;
        if (pp < npiece) v[u] = __ldcg(reinterpret_cast<const double2*>(Xg + (int64_t)(pp >> 6) * ld + 2 * (pp & 63)));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int pp = p + u * 256;
        if (pp < npiece) *reinterpret_cast<double2*>(W + (pp >> 6) * PLD + 2 * (pp & 63)) = v[u];
      }
    }
  }
  __syncthreads();
  HPROF(3);
  // ---- phase 2: A_{k+1,k+1}(own rows, column tiles <= m-tile) -= X X^T, K = 128 in four chains ----
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (tm[i] < 0) continue;
    const int jt = tj[i];
    const double* xb = W + (8 * jt + g) * PLD + q;
    const double* xa = Xs + (8 * tm[i] + g) * PLD + q;
    double2 e1 = make_double2(0.0, 0.0), e2 = e1, e3 = e1;
#pragma unroll
    for (int ks = 0; ks < NB / 4; ks += 4) {
      dmma884(c[i].x, c[i].y, -xa[4 * ks], xb[4 * ks]);
      dmma884(e1.x, e1.y, -xa[4 * ks + 4], xb[4 * ks + 4]);
      dmma884(e2.x, e2.y, -xa[4 * ks + 8], xb[4 * ks + 8]);
      dmma884(e3.x, e3.y, -xa[4 * ks + 12], xb[4 * ks + 12]);
    }
    c[i].x += (e1.x + e2.x) + e3.x; c[i].y += (e1.y + e2.y) + e3.y;
    const int row = (tm[i] ? mrow1 : mrow0) + g, col = 8 * jt + 2 * q;
    *reinterpret_cast<double2*>(Ablk + (int64_t)row * ld + (int64_t)b * NB + col) = c[i];
  }
  HPROF(4);
  cluster.sync();                                     // no CTA leaves while its rows of X are still being read
  HPROF(5);
}

// Mirrors of diagonal block kb, off the critical chain: the upper triangle of the factor block (L^T) from its lower triangle, and
// L^-1 (lower, row-major) from L^-T.  One CTA per 32 x 32 tile on or below the diagonal and per matrix (2 x 10 CTAs), transposed
// through shared memory.
__global__ void __launch_bounds__(256) chol_mirror_kernel(double* __restrict__ A, int64_t ld, int kb, double* __restrict__ Linv,
                                                          const double* __restrict__ LinvT) {
  __shared__ double T[32][33];
  const int which = blockIdx.x / 10;                   // 0: factor block, 1: inverse
  int t = blockIdx.x % 10, bi = 0;
  while (t > bi) { t -= bi + 1; ++bi; }
  const int bj = t;                                    // tile (bi, bj), bj <= bi
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (which == 0) {
    double* Ab = A + ((int64_t)kb * NB) * ld + (int64_t)kb * NB;
    for (int r = ty; r < 32; r += 8) T[r][tx] = Ab[(int64_t)(32 * bi + r) * ld + 32 * bj + tx];          // lower tile (bi, bj)
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {                 // upper tile (bj, bi): element (32 bj + r, 32 bi + tx) = L[32 bi + tx][32 bj + r]
      const int gi = 32 * bj + r, gc = 32 * bi + tx;
      if (gc > gi) Ab[(int64_t)gi * ld + gc] = T[tx][r];
    }
  } else {
    const double* LiT = LinvT + (int64_t)kb * NB * NB;
    double* Li = Linv + (int64_t)kb * NB * NB;
    for (int r = ty; r < 32; r += 8) T[r][tx] = LiT[(32 * bj + r) * NB + 32 * bi + tx];                  // upper tile (bj, bi) of W^T
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {                 // W[32 bi + r][32 bj + tx] = (W^T)[32 bj + tx][32 bi + r]
      const int gi = 32 * bi + r, gc = 32 * bj + tx;
      if (gc <= gi) Li[gi * NB + gc] = T[tx][r];
    }
  }
}

struct CholMaps { CUtensorMap L128, L64, Linv; };

// K3: one CTA per 64 rows below the diagonal block.  acc[n_panel][row] = sum_l Linv[n_panel][l] * A[row][l]  = L[row][n_panel].
__global__ void __launch_bounds__(TG_THREADS, 2) trsm_panel_kernel(double* __restrict__ A, int64_t ld, int kb, int tile0,
                                                                   const __grid_constant__ CholMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int row0 = (kb + 1) * NB + (tile0 + blockIdx.x) * TG_BN;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  tile_gemm(acc, &maps.Linv, 0, kb * NB, &maps.L64, kb * NB, row0, NB / KC, smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3, rg = rho(g);
  double* Alo = A + (int64_t)row0 * ld + (int64_t)kb * NB;          // [row][panel col]
  double* Aup = A + ((int64_t)kb * NB) * ld + row0;                 // mirrored: [panel col][row]
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int pc = wm * 32 + mt * 8 + rg, r = wn * 32 + nt * 8 + q + 4 * e;
        Alo[(int64_t)r * ld + pc] = acc[mt][nt][e];
        Aup[(int64_t)pc * ld + r] = acc[mt][nt][e];
      }
}

// K4: one CTA per 128 x 64 tile (row block bi >= bi_lo, 64-wide column block c2 in [col2_lo, min(col2_hi, 2 bi + 2)))
// of the trailing lower triangle:   A_ij -= L_i,[kb0, kb0+nkb) L_j,[kb0, kb0+nkb)^T   (K = 128 nkb).
// Two-level blocking: inside a 512-wide outer panel the update uses nkb = 1 and stops at the panel's last column; the update
// of everything to the right of the panel is ONE launch with nkb = 4, where per-tile overheads are amortised over K = 512.
__global__ void __launch_bounds__(TG_THREADS, 2) syrk_trailing_kernel(double* __restrict__ A, int64_t ld, int kb0, int nkb, int bi_lo,
                                                                      int col2_lo, int col2_hi, int nblk,
                                                                      const __grid_constant__ CholMaps maps, double* __restrict__ Cd = nullptr,
                                                                      int xbi = -1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  int t = blockIdx.x, bi = bi_lo, c2 = 0;
  for (; bi < nblk; ++bi) {
    const int hi = (2 * bi + 2 < col2_hi) ? 2 * bi + 2 : col2_hi;
    const int cnt = hi - col2_lo;
    if (cnt > 0) { if (t < cnt) { c2 = col2_lo + t; break; } t -= cnt; }
  }
  if (bi >= nblk) {
    if (xbi < 0 || t >= 2) return;
    bi = xbi; c2 = 2 * xbi + t;           // the two extra CTAs of the launch: the diagonal block of row block xbi
  }
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  tile_gemm(acc, &maps.L128, kb0 * NB, bi * NB, &maps.L64, kb0 * NB, c2 * TG_BN, nkb * (NB / KC), smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3, rg = rho(g);
  // Cd: the update of ONE diagonal block goes to a 128 x 128 scratch accumulator instead (look-ahead schedule, outer-panel boundary)
  double* C = Cd ? Cd + (int64_t)(c2 - 2 * bi) * TG_BN : A + ((int64_t)bi * NB) * ld + (int64_t)c2 * TG_BN;
  if (Cd) ld = NB;
  const int diag_off = bi * NB - c2 * TG_BN;      // element (m, n) is on/below the diagonal iff n <= m + diag_off
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int m = wm * 32 + mt * 8 + rg, n = wn * 32 + nt * 8 + q + 4 * e;
        if (n <= m + diag_off) C[(int64_t)m * ld + n] -= acc[mt][nt][e];
      }
}

static int syrk_tiles(int bi_lo, int nblk, int col2_lo, int col2_hi) {
  int n = 0;
  for (int bi = bi_lo; bi < nblk; ++bi) {
    const int hi = (2 * bi + 2 < col2_hi) ? 2 * bi + 2 : col2_hi;
    if (hi > col2_lo) n += hi - col2_lo;
  }
  return n;
}

static const int OB = (getenv("B200BO_OB") && atoi(getenv("B200BO_OB")) >= 1 && atoi(getenv("B200BO_OB")) <= 4) ? atoi(getenv("B200BO_OB")) : 4;   // inner panels per outer panel (developer knob; 4 = 512 columns)

static cudaError_t launch_cholesky_inorder(b200bo_handle_s* h) {
  const int nblk = (int)(h->Np / NB);
  const size_t sm_potrf = PD_SMEM;
  cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_potrf);
  cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM);
  cudaFuncSetAttribute(syrk_trailing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM);
  CholMaps maps;
  maps.L128 = h->tmL; maps.L64 = h->tmL64; maps.Linv = h->tmLinv;
  cudaStream_t sa = h->stream, sb = h->stream2;
  cudaMemsetAsync(h->dinfo, 0, sizeof(int), sa);
  const int npan = (nblk + OB - 1) / OB;
  while ((int)h->syrk_ev.size() < 2 * npan + 2) { cudaEvent_t e; cudaEventCreate(&e); h->syrk_ev.push_back(e); }
  while ((int)h->la_ev.size() < 2 * npan + 2) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); h->la_ev.push_back(e); }
  while ((int)h->fw_ev.size() < nblk + 2) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); h->fw_ev.push_back(e); }
  h->syrk_ev_used = 0;
  const int big = 1 << 30;
  // the forward solve z = L^-1 (y - m) rides along on stream B (solve.cu: launch_fwd_step)
  cudaEventRecord(h->fw_ev[nblk], sa);
  cudaStreamWaitEvent(sb, h->fw_ev[nblk], 0);
  launch_residual(h, sb);
  auto syrk = [&](cudaStream_t st, int kb0, int nkb, int bi_lo, int lo, int hi) {
    const int n = syrk_tiles(bi_lo, nblk, lo, hi);
    if (n > 0) {
      syrk_trailing_kernel<<<n, TG_THREADS, TG_SMEM, st>>>(h->dL, h->ld, kb0, nkb, bi_lo, lo, hi, nblk, maps);
      h->launches++;
    }
    return n;
  };
  // Outer panel P = inner panels [p0, p1).  Handle stream (A): the inner factorisation of P, then the K = 512 update of the
  // NEXT outer panel's columns; second stream (B): the K = 512 update of everything further right, overlapping panel P+1.
  for (int P = 0; P < npan; ++P) {
    const int p0 = P * OB, p1 = (p0 + OB < nblk) ? p0 + OB : nblk, p2 = (p1 + OB < nblk) ? p1 + OB : nblk;
    for (int k = p0; k < p1; ++k) {
      potrf_diag_kernel<<<1, PD_THREADS, sm_potrf, sa>>>(h->dL, h->ld, k, h->dLinv, h->dLinvT, h->dinfo);
      chol_mirror_kernel<<<20, 256, 0, sa>>>(h->dL, h->ld, k, h->dLinv, h->dLinvT);
      h->launches += 2;
      cudaEventRecord(h->fw_ev[k], sa);
      cudaStreamWaitEvent(sb, h->fw_ev[k], 0);
      launch_fwd_step(h, sb, k - 1, nblk);
      const int rem = nblk - k - 1;
      if (rem <= 0) break;
      trsm_panel_kernel<<<rem * (NB / TG_BN), TG_THREADS, TG_SMEM, sa>>>(h->dL, h->ld, k, 0, maps);
      h->launches++;
      syrk(sa, k, 1, k + 1, 2 * (k + 1), 2 * p1);            // inner update: only the columns still inside the outer panel
    }
    if (p1 >= nblk) break;
    cudaEvent_t Pk = h->la_ev[2 * P], Rk = h->la_ev[2 * P + 1];
    // K = 512 updates on tcgen05 (int8 slices, syrk_i8.cu) when enabled and the outer panel is full; else the DMMA tile GEMM
    const bool i8 = (h->syrk_engine < 0 ? syrk_i8_enabled() : h->syrk_engine >= 1) && (p1 - p0) == OB;
    if (i8) {
      if (P > 0) cudaStreamWaitEvent(sa, h->la_ev[2 * (P - 1) + 1], 0);   // the far part of panel P-1 still reads the previous slices
      launch_slice_panel(h, sa, p1 * NB, p0 * NB, p1 - p0);
    }
    // The tcgen05 kernels are persistent (one CTA per SM, 176 KB of shared memory): the far update would occupy every SM for its
    // whole duration and starve the critical chain (K2 needs a free SM).  It therefore runs on num_sms - reserve CTAs, which
    // leaves `reserve` SMs to the next outer panel's chain -- a static SM partition -- and the near part (on the chain) goes first.
    static const int reserve = getenv("B200BO_I8_RESERVE") ? atoi(getenv("B200BO_I8_RESERVE")) : 40;   // measured at N=8192: 0 -> 8.43 ms, 24 -> 7.76, 40 -> 7.51, 64 -> 7.90
    const bool have_far = syrk_tiles(p1, nblk, 2 * p2, big) > 0;
    if (have_far) { cudaEventRecord(Pk, sa); cudaStreamWaitEvent(sb, Pk, 0); }
    if (P > 0) cudaStreamWaitEvent(sa, h->la_ev[2 * (P - 1) + 1], 0);   // far part of panel P-1 also wrote the next panel's columns
    if (i8) launch_syrk_i8(h, sa, p1, 2 * p1, 2 * p2, nullptr, h->num_sms); else syrk(sa, p0, p1 - p0, p1, 2 * p1, 2 * p2);   // near part: the next outer panel's columns
    if (have_far) {                                           // far part on stream B (after panel P is complete on A)
      cudaEventRecord(h->syrk_ev[h->syrk_ev_used++], sb);
      if (i8) launch_syrk_i8(h, sb, p1, 2 * p2, big, nullptr, std::max(8, h->num_sms - reserve)); else syrk(sb, p0, p1 - p0, p1, 2 * p2, big);
      cudaEventRecord(h->syrk_ev[h->syrk_ev_used++], sb);
      cudaEventRecord(Rk, sb);
    }
  }
  cudaEventRecord(h->fw_ev[nblk + 1], sb);
  cudaStreamWaitEvent(sa, h->fw_ev[nblk + 1], 0);                             // join stream B (forward solve and far updates)
#ifdef POTRF_PROF
  {
    cudaDeviceSynchronize();
    unsigned long long pr[16];
    cudaMemcpyFromSymbol(pr, g_potrf_prof, sizeof(pr));
    fprintf(stderr, "potrf phases (cycles per kernel, %d kernels): load %llu | F+U2 %llu | T+inv %llu | U1 %llu | offdiag inverse %llu | write-back %llu\n", nblk,
            pr[0] / nblk, pr[1] / nblk, pr[2] / nblk, pr[3] / nblk, pr[4] / nblk, pr[5] / nblk);
    fprintf(stderr, "  F(s) alone on warp 0: %llu %llu %llu %llu | phase A(s): %llu %llu %llu %llu\n", pr[8] / nblk, pr[9] / nblk, pr[10] / nblk, pr[11] / nblk,
            pr[12] / nblk, pr[13] / nblk, pr[14] / nblk, pr[15] / nblk);
    memset(pr, 0, sizeof(pr));
    cudaMemcpyToSymbol(g_potrf_prof, pr, sizeof(pr));
  }
#endif
  return cudaGetLastError();
}

// Look-ahead schedule (round 2).  The only work between two diagonal blocks that the NEXT diagonal block needs is the "head" of panel k:
// L_{k+1,k} (one 128-row slab of the panel solve) and the rank-128 update of A_{k+1,k+1}.  The chain stream therefore runs
//     potrf(k) -> head(k) -> potrf(k+1) -> ...
// and everything else of panel k runs underneath potrf(k+1), itself split by what the next head needs:
//   stream C (column path)  R1(k): the panel solve below block k+1;  R2a(k): the update of block column k+1 and of the diagonal block
//                           A_{k+2,k+2} -- what head(k+1) and R1(k+1) read -- then event Re(k);  at an outer-panel boundary the int8
//                           slicing and the near K = 512 update (tcgen05);
//   stream D (bulk)         R2b(k): the rest of the inner update (columns k+2 .. end of the outer panel);
//   stream E (riders)       the forward solve of y - m, and panel k's share of the boundary block A_{p1,p1} into a scratch accumulator
//                           (A_{p1,p1} itself is still being written by the far update of the previous outer panel);
//   stream B (far)          the far K = 512 update, low priority, on num_sms - reserve SMs as before.
// Every tile still receives its updates in a fixed order (events), so repeated factorisations stay bit-identical.
static cudaError_t launch_cholesky_lookahead(b200bo_handle_s* h, bool fused_head, bool capturing) {
  const int nblk = (int)(h->Np / NB);
  cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PD_SMEM);
  cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM);
  cudaFuncSetAttribute(syrk_trailing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM);
  cudaFuncSetAttribute(chol_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM);
  CholMaps maps;
  maps.L128 = h->tmL; maps.L64 = h->tmL64; maps.Linv = h->tmLinv;
  if (!h->dD) { const cudaError_t e = cudaMalloc(&h->dD, sizeof(double) * NB * NB); if (e != cudaSuccess) return e; }
  if (!h->stream3) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const int mid = hi < lo ? hi + 1 : hi;
    cudaError_t e = cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, mid);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->stream4, cudaStreamNonBlocking, mid < lo ? mid + 1 : mid);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->stream5, cudaStreamNonBlocking, mid < lo ? mid + 1 : mid);
    if (e != cudaSuccess) return e;
  }
  cudaStream_t sa = h->stream, sb = h->stream2, sc = h->stream3, sd = h->stream4, se = h->stream5;
  cudaMemsetAsync(h->dinfo, 0, sizeof(int), sa);
  const int npan = (nblk + OB - 1) / OB;
  auto grow = [](std::vector<cudaEvent_t>& v, size_t n, bool timing) {
    while (v.size() < n) { cudaEvent_t e; if (timing) cudaEventCreate(&e); else cudaEventCreateWithFlags(&e, cudaEventDisableTiming); v.push_back(e); }
  };
  grow(h->syrk_ev, 2 * npan + 2, true);
  grow(h->la_ev, 3 * npan + 2, false);
  grow(h->fw_ev, nblk + 2, false);
  static const bool trace_env = getenv("B200BO_CHOL_TRACE") != nullptr;  // developer knob: per-panel event timeline on stderr (eager launches only)
  const bool trace = trace_env && !capturing;
  grow(h->ch_ev, 6 * (size_t)nblk + 8, trace);
  h->syrk_ev_used = 0;
  const int big = 1 << 30;
  auto Pe = [&](int k) { return h->ch_ev[6 * k]; };        // potrf(k) done (chain)
  auto He = [&](int k) { return h->ch_ev[6 * k + 1]; };    // head(k) done (chain)
  auto Re = [&](int k) { return h->ch_ev[6 * k + 2]; };    // column path of panel k done (C)
  auto R1e = [&](int k) { return h->ch_ev[6 * k + 3]; };   // panel solve of panel k done (C)
  auto Be = [&](int k) { return h->ch_ev[6 * k + 4]; };    // bulk inner update of panel k done (D)
  auto De = [&](int k) { return h->ch_ev[6 * k + 5]; };    // panel k's share of the boundary block accumulated (E)
  auto syrk = [&](cudaStream_t st, int kb0, int nkb, int bi_lo, int bi_hi, int lo, int hi, double* Cd = nullptr, int xbi = -1) {
    const int n = syrk_tiles(bi_lo, bi_hi, lo, hi) + (xbi >= 0 ? 2 : 0);
    if (n > 0) {
      syrk_trailing_kernel<<<n, TG_THREADS, TG_SMEM, st>>>(h->dL, h->ld, kb0, nkb, bi_lo, lo, hi, bi_hi, maps, Cd, xbi);
      h->launches++;
    }
    return n;
  };
  auto trsm = [&](cudaStream_t st, int k, int tile0, int ntile) {
    if (ntile > 0) {
      trsm_panel_kernel<<<ntile, TG_THREADS, TG_SMEM, st>>>(h->dL, h->ld, k, tile0, maps);
      h->launches++;
    }
  };
  static const int reserve = getenv("B200BO_I8_RESERVE") ? atoi(getenv("B200BO_I8_RESERVE")) : 40;
  static const int near_reserve = getenv("B200BO_I8_NEAR_RESERVE") ? atoi(getenv("B200BO_I8_NEAR_RESERVE")) : 32;
  // start / stop of the far updates (B200BO_T_SYRK): inside a captured graph they must be event-record NODES
  auto stamp = [&](cudaEvent_t e, cudaStream_t st) { if (capturing) cudaEventRecordWithFlags(e, st, cudaEventRecordExternal); else cudaEventRecord(e, st); };
  cudaEventRecord(h->fw_ev[nblk], sa);
  if (trace) cudaEventRecord(h->syrk_ev[2 * npan + 1], sa);
  for (cudaStream_t st : {sb, sc, sd, se}) cudaStreamWaitEvent(st, h->fw_ev[nblk], 0);
  // (the forward solve z = L^-1 (y - m) rides along on the rider stream -- solve.cu: launch_fwd_step; its right-hand side y - m was
  //  written by launch_cholesky BEFORE this schedule, outside any captured graph: it depends on beta and N, which a graph would freeze)
  for (int P = 0; P < npan; ++P) {
    const int p0 = P * OB, p1 = (p0 + OB < nblk) ? p0 + OB : nblk, p2 = (p1 + OB < nblk) ? p1 + OB : nblk;
    const bool use_D = fused_head && p1 < nblk && p1 - p0 > 1;
    for (int k = p0; k < p1; ++k) {
      potrf_diag_kernel<<<1, PD_THREADS, PD_SMEM, sa>>>(h->dL, h->ld, k, h->dLinv, h->dLinvT, h->dinfo);
      h->launches++;
      // the transposed triangles of the block (L^T mirror, row-major L^-1) are nobody's business on the chain: the fused head reads L^-T
      if (!fused_head) { chol_mirror_kernel<<<20, 256, 0, sa>>>(h->dL, h->ld, k, h->dLinv, h->dLinvT); h->launches++; }
      cudaEventRecord(Pe(k), sa);
      if (fused_head) {
        cudaStreamWaitEvent(sc, Pe(k), 0);
        chol_mirror_kernel<<<20, 256, 0, sc>>>(h->dL, h->ld, k, h->dLinv, h->dLinvT);
        h->launches++;
      }
      cudaStreamWaitEvent(se, Pe(k), 0);
      if (use_D && k == p0) cudaMemsetAsync(h->dD, 0, sizeof(double) * NB * NB, se);   // its last reader, head(p0-1), precedes potrf(p0)
      if (k > 0) cudaStreamWaitEvent(se, R1e(k - 1), 0);                    // L_{.,k-1} complete
      launch_fwd_step(h, se, k - 1, nblk);
      const int rem = nblk - k - 1;
      if (rem <= 0) break;
      // ---- head(k) on the chain ----
      const bool boundary = k + 1 == p1;                                    // A_{p1,p1} lies outside the inner updates
      if (k > 0) cudaStreamWaitEvent(sa, Re(k - 1), 0);                     // A_{k+1,k}, A_{k+1,k+1} carry every update of the panels < k
      if (boundary && P > 0) cudaStreamWaitEvent(sa, h->la_ev[3 * (P - 1) + 1], 0);   // the far update of panel P-1 also writes A_{p1,p1}
      if (fused_head) {
        if (boundary && k > p0) cudaStreamWaitEvent(sa, De(k - 1), 0);      // panels p0 .. p1-2 left their share in the scratch accumulator
        chol_head_kernel<<<HC, 256, HEAD_SMEM, sa>>>(h->dL, h->ld, k, h->dLinvT, (boundary && k > p0) ? h->dD : nullptr);
        h->launches++;
      } else {
        trsm(sa, k, 0, NB / TG_BN);                                        // L_{k+1,k}
        if (!boundary) syrk(sa, k, 1, k + 1, k + 2, 2 * (k + 1), 2 * (k + 1) + 2);   // A_{k+1,k+1} -= L_{k+1,k} L_{k+1,k}^T
        else syrk(sa, p0, p1 - p0, p1, p1 + 1, 2 * p1, 2 * p1 + 2);        // outer boundary: the whole K = 128 (p1 - p0) update of A_{p1,p1}
      }
      cudaEventRecord(He(k), sa);
      // ---- column path of panel k (stream C) ----
      cudaStreamWaitEvent(sc, Pe(k), 0);
      trsm(sc, k, NB / TG_BN, (rem - 1) * (NB / TG_BN));                   // R1: L_{i,k}, i >= k+2
      cudaEventRecord(R1e(k), sc);
      cudaStreamWaitEvent(sc, He(k), 0);                                    // L_{k+1,k}
      if (k > p0) cudaStreamWaitEvent(sc, Be(k - 1), 0);                    // the bulk update of panel k-1 wrote these tiles first
      if (!boundary) syrk(sc, k, 1, k + 2, nblk, 2 * (k + 1), 2 * (k + 1) + 2, nullptr, (k + 2 < p1) ? k + 2 : -1);   // R2a: block column k+1 (+ A_{k+2,k+2})
      if (!boundary) cudaEventRecord(Re(k), sc);
      // ---- bulk inner update (stream D): columns k+2 .. p1-1 below the diagonal blocks taken by R2a ----
      if (k + 2 < p1) {
        cudaStreamWaitEvent(sd, R1e(k), 0);
        syrk(sd, k, 1, k + 3, nblk, 2 * (k + 2), 2 * p1);
      }
      cudaEventRecord(Be(k), sd);
      // ---- this panel's share of the boundary block, off the chain (stream E) ----
      if (use_D && !boundary) {
        cudaStreamWaitEvent(se, R1e(k), 0);
        syrk(se, k, 1, p1, p1 + 1, 2 * p1, 2 * p1 + 2, h->dD);
        cudaEventRecord(De(k), se);
      }
    }
    if (p1 >= nblk) break;
    // ---- outer panel P is complete on stream C: K = 512 updates (tcgen05 int8 slices when enabled and the panel is full) ----
    cudaEvent_t Sk = h->la_ev[3 * P], Rk = h->la_ev[3 * P + 1];
    const bool i8 = (h->syrk_engine < 0 ? syrk_i8_enabled() : h->syrk_engine >= 1) && (p1 - p0) == OB;
    if (P > 0) cudaStreamWaitEvent(sc, h->la_ev[3 * (P - 1) + 1], 0);       // far(P-1) still reads the previous slices and writes these columns
    if (i8) launch_slice_panel(h, sc, p1 * NB, p0 * NB, p1 - p0);
    const bool have_far = syrk_tiles(p1, nblk, 2 * p2, big) > 0;
    if (have_far) { cudaEventRecord(Sk, sc); cudaStreamWaitEvent(sb, Sk, 0); }
    // near part: the next outer panel's columns below its first diagonal block (head(p1-1) took A_{p1,p1})
    if (i8) launch_syrk_i8(h, sc, p1 + 1, 2 * p1, 2 * p2, nullptr, std::max(8, h->num_sms - near_reserve));
    else syrk(sc, p0, p1 - p0, p1 + 1, nblk, 2 * p1, 2 * p2);
    cudaEventRecord(Re(p1 - 1), sc);
    if (have_far) {
      stamp(h->syrk_ev[h->syrk_ev_used++], sb);
      if (i8) launch_syrk_i8(h, sb, p1, 2 * p2, big, nullptr, std::max(8, h->num_sms - reserve)); else syrk(sb, p0, p1 - p0, p1, nblk, 2 * p2, big);
      stamp(h->syrk_ev[h->syrk_ev_used++], sb);
      cudaEventRecord(Rk, sb);
    }
  }
#ifdef POTRF_PROF
  if (!capturing) {
    cudaDeviceSynchronize();
    unsigned long long hp[8];
    cudaMemcpyFromSymbol(hp, g_head_prof, sizeof(hp));
    const int n = nblk > 1 ? nblk - 1 : 1;
    fprintf(stderr, "head phases (cycles per kernel): load wait %llu | phase 1 %llu | sync %llu | gather %llu | phase 2 %llu | sync %llu\n", hp[0] / n, hp[1] / n, hp[2] / n,
            hp[3] / n, hp[4] / n, hp[5] / n);
    memset(hp, 0, sizeof(hp));
    cudaMemcpyToSymbol(g_head_prof, hp, sizeof(hp));
  }
#endif
  if (trace) {
    cudaEventRecord(h->ch_ev[6 * nblk + 5], sa);
    cudaStreamSynchronize(sa); cudaStreamSynchronize(sb); cudaStreamSynchronize(sc); cudaStreamSynchronize(sd); cudaStreamSynchronize(se);
    auto at = [&](cudaEvent_t e) { float ms = -1.f; if (cudaEventElapsedTime(&ms, h->syrk_ev[2 * npan + 1], e) != cudaSuccess) { cudaGetLastError(); return -1.0; } return (double)ms * 1e3; };
    fprintf(stderr, "# k: potrf done | head done | panel solve done | column path done | bulk done   (us since start)\n");
    for (int k = 0; k < nblk - 1; ++k)
      fprintf(stderr, "%3d: %9.1f %9.1f %9.1f %9.1f %9.1f\n", k, at(Pe(k)), at(He(k)), at(R1e(k)), at(Re(k)), at(Be(k)));
    fprintf(stderr, "chain end %9.1f\n", at(h->ch_ev[6 * nblk + 5]));
  }
  int j = 0;
  for (cudaStream_t st : {sb, sc, sd, se}) {               // join
    cudaEventRecord(h->ch_ev[6 * nblk + j], st);
    cudaStreamWaitEvent(sa, h->ch_ev[6 * nblk + j], 0);
    ++j;
  }
  return cudaGetLastError();
}

// The look-ahead schedule is ~25 launches and ~40 event operations per 128-column panel over five streams: enqueued one by one the HOST
// becomes the critical path (the chain itself is ~50 us per panel).  The whole factorisation is therefore captured ONCE per shape into
// a CUDA graph and replayed with one launch.
// The graph bakes in pointers and panel counts: it is keyed on them and rebuilt when they change.
static void drop_chol_graph(b200bo_handle_s* h) {
  if (h->chol_graph_exec) cudaGraphExecDestroy(h->chol_graph_exec);
  h->chol_graph_exec = nullptr; h->chol_graph_key = 0;
}

void release_cholesky_graph(b200bo_handle_s* h) { drop_chol_graph(h); }

cudaError_t launch_cholesky(b200bo_handle_s* h) {
  static const int sched = getenv("B200BO_CHOL_SCHED") ? atoi(getenv("B200BO_CHOL_SCHED")) : 1;   // developer knob: 0 = the in-order schedule of round 1
  static const int graph_env = getenv("B200BO_CHOL_GRAPH") ? atoi(getenv("B200BO_CHOL_GRAPH")) : 1;
  const int sc = h->chol_sched >= 0 ? h->chol_sched : sched;       // 1: look-ahead with the fused cluster head, 2: look-ahead with tile-GEMM heads
  if (!sc) return launch_cholesky_inorder(h);
  // dw = y - m: the ONE launch of the factorisation whose arguments change with the data (N) and the parameters (beta) while the panel
  // count and the buffers -- the graph's key -- stay the same.  It runs eagerly, ahead of the (possibly replayed) schedule.
  { const cudaError_t er = launch_residual(h, h->stream); if (er != cudaSuccess) return er; }
  const bool want_graph = (h->chol_graph >= 0 ? h->chol_graph : graph_env) != 0;
  if (!want_graph) return launch_cholesky_lookahead(h, sc == 1, false);
  const bool i8 = h->syrk_engine < 0 ? syrk_i8_enabled() : h->syrk_engine >= 1;
  uint64_t key = 1469598103934665603ull;
  for (uint64_t v : {(uint64_t)(h->Np / NB), (uint64_t)sc, (uint64_t)(i8 ? 1 + h->syrk_engine : 0), (uint64_t)(uintptr_t)h->dL, (uint64_t)(uintptr_t)h->dLinv,
                     (uint64_t)(uintptr_t)h->dSl, (uint64_t)(uintptr_t)h->dw, (uint64_t)h->ld, (uint64_t)(uintptr_t)h->stream})
    key = (key ^ v) * 1099511628211ull;
  if (h->chol_graph_exec && h->chol_graph_key == key) {
    h->launches += h->chol_graph_launches;
    h->syrk_ev_used = h->chol_graph_syrk_ev;
    return cudaGraphLaunch(h->chol_graph_exec, h->stream);
  }
  // Capturing and instantiating the ~25 nblk nodes is not free (measured: 3.3 ms at N = 2048, 8 ms at N = 4096, 17 ms at N = 8192) against
  // 0.12-0.35 ms saved per replay: a shape is captured at its `thr`-th consecutive factorisation (default 6; knob "chol_graph" = k >= 2
  // sets it), so that a one-off refit never pays and the evaluations of a MAP fit do from the sixth on.  The first sight is always eager
  // (it makes every lazy allocation).
  const int knob = h->chol_graph >= 0 ? h->chol_graph : graph_env;
  const int thr = knob <= 1 ? 6 : knob;
  h->chol_seen_count = h->chol_seen_key == key ? h->chol_seen_count + 1 : 1;
  h->chol_seen_key = key;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (h->chol_seen_count < thr || cudaStreamIsCapturing(h->stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone)
    return launch_cholesky_lookahead(h, sc == 1, false);             // (also: a caller stream that is itself capturing)
  drop_chol_graph(h);
  const int64_t launches0 = h->launches;
  cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) { cudaGetLastError(); return launch_cholesky_lookahead(h, sc == 1, false); }
  e = launch_cholesky_lookahead(h, sc == 1, true);
  cudaGraph_t graph = nullptr;
  const cudaError_t e2 = cudaStreamEndCapture(h->stream, &graph);
  if (e != cudaSuccess || e2 != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    h->launches = launches0;
    h->chol_graph = 0;                                               // capture is not possible here: stay eager
    return launch_cholesky_lookahead(h, sc == 1, false);
  }
  e = cudaGraphInstantiate(&h->chol_graph_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { cudaGetLastError(); h->chol_graph_exec = nullptr; h->launches = launches0; h->chol_graph = 0; return launch_cholesky_lookahead(h, sc == 1, false); }
  h->chol_graph_key = key;
  h->chol_graph_launches = h->launches - launches0;
  h->chol_graph_syrk_ev = h->syrk_ev_used;
  return cudaGraphLaunch(h->chol_graph_exec, h->stream);
}

}  // namespace b200bo
