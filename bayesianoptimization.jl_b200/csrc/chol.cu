// chol.cu -- K2-K4: blocked right-looking Cholesky of Sigma (FP64), in place on the lower triangle of h->dL.
//
// Replaces EXT `cholesky!(Symmetric(.., :U))` inside GaussianProcesses.jl `make_posdef!` (reached from the
// reference's update!(model, x, y), src/models/gp.jl:11-18).  The reference factor is the upper U with
// Sigma = U'U in column-major storage; that is byte-identical to the row-major lower L = U' kept here.
//
//   for each 128-wide panel k:
//     K2 potrf_diag   : L_kk = chol(A_kk) by ONE 1024-thread CTA, one barrier per column; the same sweep carries the
//                       forward substitution on I, so L_kk^-1 (needed by every GEMM-shaped solve) costs no extra pass.
//     K3 trsm_panel   : A_ik <- A_ik L_kk^-T = (L_kk^-1 A_ik^T)^T                (TMA + DMMA tile GEMM, K = 128)
//     K4 syrk_trailing: A_ij <- A_ij - L_ik L_jk^T  for i >= j > k              (TMA + DMMA tile GEMM: the dense contraction)
//   The upper triangle of h->dL receives L^T ("mirrored factor") so the backward solve L^T w = v reads the same
//   k-major rows as the forward one.
//   Two-level blocking: inner panels of 128 inside outer panels of 512; the update right of an outer panel is one K = 512
//   launch.  Look-ahead: the next outer panel is factorised on the handle's stream while the far part of the previous
//   K = 512 update runs on a second stream.
#include "tma.cuh"
#include "handle.h"

namespace b200bo {

// ------------------------------------------------------------------------------------------------------------
// K2: diagonal block, register resident, blocked by 4 columns.  One thread per 4 x 4 tile of the lower triangle (528 of 544
// threads); thread (ti, tk) keeps rows 4 ti + b, cols 4 tk + a in registers: on and below the diagonal it starts as A and, panel
// by panel, turns into B = the forward substitution applied to I (L^-1 with unscaled rows).  Per 4-column panel j4:
//   (a) the diagonal tile's thread factors its 4 x 4 block locally and publishes L44 and the reciprocal pivots;   -- barrier
//   (b) the tile-column owners (ti > j4, tk = j4) solve X L44^T = A for their 4 x 4 blocks and publish the 4 finished columns
//       of L; the tile-row owners (ti = j4) finish their 4 rows of B locally and publish them;                      -- barrier
//   (c) every tile below the panel applies a rank-4 update: A -= Lcol Lcol^T (tk > j4) or B -= (Lcol / piv) Brow (tk <= j4).
// Two barriers and ~110 instructions per thread per FOUR columns; no shared-memory read-modify-write.  Positions above the
// diagonal collect garbage and are never read.
// ------------------------------------------------------------------------------------------------------------
constexpr int PS = NB + 1;
constexpr int PD_THREADS = 544;   // 17 warps >= 32*33/2 = 528 lower-triangle tiles

__global__ void __launch_bounds__(PD_THREADS, 1) potrf_diag_kernel(double* __restrict__ A, int64_t ld, int kb, double* __restrict__ Linv,
                                                                   double* __restrict__ LinvT, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* Lf = sm;                       // [NB][PS] finished columns of L (for the coalesced write-back)
  double* colbuf = Lf + NB * PS;         // [4][NB] the panel's finished columns of L        (NB*PS is even: 16-byte aligned)
  double* rowbuf = colbuf + 4 * NB;      // [4][NB] the panel's finished rows of B
  double* l44 = rowbuf + 4 * NB;         // [2][16] the panel's diagonal block of L (row-major, strictly lower used)
  double* rinv4 = l44 + 32;              // [2][4]  reciprocal pivots
  double* dg = rinv4 + 8;                // [NB] pivots (raw until the end, then sqrt)
  const int tid = threadIdx.x;
  int ti = (int)((sqrt(8.0 * (double)tid + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= tid) ++ti;
  while (ti * (ti + 1) / 2 > tid) --ti;
  int tk = tid - ti * (ti + 1) / 2;
  const bool active = tid < 528;                     // tile (ti, tk), tk <= ti, of the lower triangle
  if (!active) { ti = 64; tk = 64; }                 // matches no panel: idle threads only join the barriers
  double* Ab = A + ((int64_t)kb * NB) * ld + (int64_t)kb * NB;
  double r[4][4];
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = 4 * ti + b, k = 4 * tk + a;
      r[b][a] = (active && k <= i) ? Ab[(int64_t)i * ld + k] : 0.0;
    }
#pragma unroll 1
  for (int j4 = 0; j4 < NB / 4; ++j4) {
    const int par = j4 & 1;
    double* L4 = l44 + par * 16;
    double* R4 = rinv4 + par * 4;
    // ---------------- (a) diagonal tile: local 4 x 4 Cholesky, its block of B, its rows of rowbuf ----------------
    if (ti == j4 && tk == j4) {
      double rs[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double d = r[c][c];
        if (!(d > 0.0)) { atomicCAS(info, 0, kb * NB + 4 * j4 + c + 1); d = 1.0; }
        dg[4 * j4 + c] = d;
        rs[c] = rsqrt(d);
        R4[c] = rs[c];
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (b > c) r[b][c] *= rs[c];                         // l_{b,c}
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a)
            if (b > c && a > c && a <= b) r[b][a] = fma(-r[b][c], r[a][c], r[b][a]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (b > c) { L4[b * 4 + c] = r[b][c]; Lf[(4 * j4 + b) * PS + 4 * j4 + c] = r[b][c]; }
      // unit-lower block of B:  B[b][a] = -( f(b,a) + sum_{l=a+1}^{b-1} f(b,l) B[l][a] ),  f(b,l) = l_{b,l} / l_{l,l}
      double bt[4][4];
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a) bt[b][a] = 0.0;
#pragma unroll
      for (int b = 1; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (a < b) {
            double sacc = r[b][a] * rs[a];
#pragma unroll
            for (int l = 1; l < 4; ++l)
              if (l > a && l < b) sacc = fma(r[b][l] * rs[l], bt[l][a], sacc);
            bt[b][a] = -sacc;
          }
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          r[b][a] = (a < b) ? bt[b][a] : 0.0;
          rowbuf[b * NB + 4 * j4 + a] = (a < b) ? bt[b][a] : (a == b ? 1.0 : 0.0);
        }
    }
    __syncthreads();
    // ---------------- (b) tile-column owners solve, tile-row owners finish their rows of B ----------------
    if (tk == j4 && ti > j4) {
      double x[4][4];
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double v = r[b][c];
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2)
            if (c2 < c) v = fma(-x[b][c2], L4[c * 4 + c2], v);
          x[b][c] = v * R4[c];
        }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        *reinterpret_cast<double2*>(colbuf + c * NB + 4 * ti) = make_double2(x[0][c], x[1][c]);
        *reinterpret_cast<double2*>(colbuf + c * NB + 4 * ti + 2) = make_double2(x[2][c], x[3][c]);
#pragma unroll
        for (int b = 0; b < 4; ++b) { Lf[(4 * ti + b) * PS + 4 * j4 + c] = x[b][c]; r[b][c] = 0.0; }
      }
    } else if (ti == j4 && tk < j4) {
#pragma unroll
      for (int b = 1; b < 4; ++b)
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (l < b) {
            const double f = L4[b * 4 + l] * R4[l];
#pragma unroll
            for (int a = 0; a < 4; ++a) r[b][a] = fma(-f, r[l][a], r[b][a]);
          }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        *reinterpret_cast<double2*>(rowbuf + b * NB + 4 * tk) = make_double2(r[b][0], r[b][1]);
        *reinterpret_cast<double2*>(rowbuf + b * NB + 4 * tk + 2) = make_double2(r[b][2], r[b][3]);
      }
    }
    __syncthreads();
    // ---------------- (c) rank-4 update of every tile below the panel ----------------
    if (active && ti > j4) {
      double li[4][4];                                       // li[c][b] = L[4 ti + b][4 j4 + c]
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 u = *reinterpret_cast<const double2*>(colbuf + c * NB + 4 * ti);
        const double2 v = *reinterpret_cast<const double2*>(colbuf + c * NB + 4 * ti + 2);
        li[c][0] = u.x; li[c][1] = u.y; li[c][2] = v.x; li[c][3] = v.y;
      }
      if (tk > j4) {                                         // Cholesky: A -= Lcol Lcol^T
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double2 u = *reinterpret_cast<const double2*>(colbuf + c * NB + 4 * tk);
          const double2 v = *reinterpret_cast<const double2*>(colbuf + c * NB + 4 * tk + 2);
          const double lk[4] = {u.x, u.y, v.x, v.y};
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int a = 0; a < 4; ++a) r[b][a] = fma(-li[c][b], lk[a], r[b][a]);
        }
      } else {                                               // forward substitution on I: B -= (Lcol / piv) Brow
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double rc = R4[c];
          const double2 u = *reinterpret_cast<const double2*>(rowbuf + c * NB + 4 * tk);
          const double2 v = *reinterpret_cast<const double2*>(rowbuf + c * NB + 4 * tk + 2);
          const double bj[4] = {u.x, u.y, v.x, v.y};
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const double f = li[c][b] * rc;
#pragma unroll
            for (int a = 0; a < 4; ++a) r[b][a] = fma(-f, bj[a], r[b][a]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (tid < NB) dg[tid] = sqrt(dg[tid]);
  // the last column (jj = NB - 1) has no rows below it: nothing left for phase 2
  __syncthreads();
  // ---- L^-1 and its transpose straight from the registers: Linv[i][k] = B[i][k] / l_ii ----
  double* Li = Linv + (int64_t)kb * NB * NB;
  double* LiT = LinvT + (int64_t)kb * NB * NB;
  if (active) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = 4 * ti + b;
      const double inv_i = 1.0 / dg[i];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int k = 4 * tk + a;
        const double y = (k < i) ? r[b][a] * inv_i : (k == i ? inv_i : 0.0);
        Li[i * NB + k] = y;
        LiT[k * NB + i] = y;
        if (ti != tk) { Li[k * NB + i] = 0.0; LiT[i * NB + k] = 0.0; }     // the strictly upper tile of L^-1 is zero
      }
    }
  }
  // ---- mirrored factor block ----
  for (int e = tid; e < NB * NB; e += PD_THREADS) {
    const int i = e >> 7, c = e & 127;
    const int lo = i < c ? i : c, hi = i < c ? c : i;
    Ab[(int64_t)i * ld + c] = (i == c) ? dg[i] : Lf[hi * PS + lo];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Tile GEMM core shared by K3 and K4:  acc(128 x 64) = A(128 rows x 128 k) * B(64 rows x 128 k)^T, operands k-major in
// global memory, staged by TMA (128B swizzle) through a 4-deep ring refilled once; 8 warps of DMMA.8x8x4.
// acc[mt][nt][e] <-> (A row wm*32 + 8 mt + rho(g),  B row wn*32 + 8 nt + q + 4 e).
// ------------------------------------------------------------------------------------------------------------
constexpr int TG_STAGES = 4, TG_THREADS = 256, TG_BM = 128, TG_BN = 64;
constexpr int TG_STAGE_DBL = (TG_BM + TG_BN) * KC;
constexpr uint32_t TG_A_BYTES = TG_BM * KC * 8, TG_B_BYTES = TG_BN * KC * 8;
constexpr size_t TG_SMEM = (size_t)TG_STAGES * TG_STAGE_DBL * 8 + 64;

__device__ __forceinline__ void tile_gemm(double (&acc)[4][4][2], const CUtensorMap* mapA, int ak0, int arow, const CUtensorMap* mapB,
                                          int bk0, int brow, int nch, uint8_t* smem_raw) {
  double* stages = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + TG_STAGES * TG_STAGE_DBL);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3, rg = rho(g);
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  if (tid == 0) {
    for (int s = 0; s < TG_STAGES; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int c) {      // chunk c -> stage c % 4
    const int s = c & (TG_STAGES - 1);
    mbar_arrive_expect_tx(&full[s], TG_A_BYTES + TG_B_BYTES);
    tma_load_2d(stages + s * TG_STAGE_DBL, mapA, &full[s], ak0 + c * KC, arow);
    tma_load_2d(stages + s * TG_STAGE_DBL + TG_BM * KC, mapB, &full[s], bk0 + c * KC, brow);
  };
  if (tid == 0)
    for (int c = 0; c < TG_STAGES && c < nch; ++c) issue(c);
  const FragAddr fa(rg, q);
  const uint32_t stage0 = smem_u32(stages);
  const uint32_t a_row = (uint32_t)(wm * 32 + rg) * 128u, b_row = (uint32_t)(wn * 32 + rg) * 128u;
  // chunks are consumed in pairs; after a pair every warp meets once and thread 0 refills both stages (nch is even)
#pragma unroll 1
  for (int c = 0; c < nch; c += 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int s = (c + h) & (TG_STAGES - 1);
      mbar_wait(&full[s], (uint32_t)((c + h) / TG_STAGES) & 1u);
      const uint32_t st = stage0 + (uint32_t)s * (TG_STAGE_DBL * 8);
      warp_mma_chunk_t<4, 4, 128, 128>(acc, st + a_row, st + TG_BM * KC * 8 + b_row, fa);
    }
    if (c + TG_STAGES < nch) {
      __syncthreads();                               // every warp is done with this pair of stages before the refill
      if (tid == 0) { issue(c + TG_STAGES); issue(c + TG_STAGES + 1); }
    }
  }
}

struct CholMaps { CUtensorMap L128, L64, Linv; };

// K3: one CTA per 64 rows below the diagonal block.  acc[n_panel][row] = sum_l Linv[n_panel][l] * A[row][l]  = L[row][n_panel].
__global__ void __launch_bounds__(TG_THREADS, 2) trsm_panel_kernel(double* __restrict__ A, int64_t ld, int kb,
                                                                   const __grid_constant__ CholMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int row0 = (kb + 1) * NB + blockIdx.x * TG_BN;
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
                                                                      const __grid_constant__ CholMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  int t = blockIdx.x, bi = bi_lo, c2 = 0;
  for (; bi < nblk; ++bi) {
    const int hi = (2 * bi + 2 < col2_hi) ? 2 * bi + 2 : col2_hi;
    const int cnt = hi - col2_lo;
    if (cnt > 0) { if (t < cnt) { c2 = col2_lo + t; break; } t -= cnt; }
  }
  if (bi >= nblk) return;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  tile_gemm(acc, &maps.L128, kb0 * NB, bi * NB, &maps.L64, kb0 * NB, c2 * TG_BN, nkb * (NB / KC), smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3, rg = rho(g);
  double* C = A + ((int64_t)bi * NB) * ld + (int64_t)c2 * TG_BN;
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

constexpr int OB = 4;   // outer panel = 4 inner panels = 512 columns

cudaError_t launch_cholesky(b200bo_handle_s* h) {
  const int nblk = (int)(h->Np / NB);
  const size_t sm_potrf = (size_t)(NB * PS + 4 * NB + 4 * NB + 32 + 8 + NB) * sizeof(double);
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
  h->syrk_ev_used = 0;
  const int big = 1 << 30;
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
      h->launches++;
      const int rem = nblk - k - 1;
      if (rem <= 0) break;
      trsm_panel_kernel<<<rem * (NB / TG_BN), TG_THREADS, TG_SMEM, sa>>>(h->dL, h->ld, k, maps);
      h->launches++;
      syrk(sa, k, 1, k + 1, 2 * (k + 1), 2 * p1);            // inner update: only the columns still inside the outer panel
    }
    if (p1 >= nblk) break;
    cudaEvent_t Pk = h->la_ev[2 * P], Rk = h->la_ev[2 * P + 1];
    if (syrk_tiles(p1, nblk, 2 * p2, big) > 0) {              // far part on stream B (after panel P is complete on A)
      cudaEventRecord(Pk, sa);
      cudaStreamWaitEvent(sb, Pk, 0);
      cudaEventRecord(h->syrk_ev[h->syrk_ev_used++], sb);
      syrk(sb, p0, p1 - p0, p1, 2 * p2, big);
      cudaEventRecord(h->syrk_ev[h->syrk_ev_used++], sb);
      cudaEventRecord(Rk, sb);
    }
    if (P > 0) cudaStreamWaitEvent(sa, h->la_ev[2 * (P - 1) + 1], 0);   // far part of panel P-1 also wrote the next panel's columns
    syrk(sa, p0, p1 - p0, p1, 2 * p1, 2 * p2);                // near part: the next outer panel's columns
  }
  if (npan >= 2) cudaStreamWaitEvent(sa, h->la_ev[2 * (npan - 2) + 1], 0);   // join stream B
  return cudaGetLastError();
}

}  // namespace b200bo
