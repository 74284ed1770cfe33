// acq.cu -- K6: the fused batched acquisition step (one launch over all M candidate columns).
//
// Replaces, per BO iteration, the reference's multi-restart search: acquire_max (src/acquisition.jl:54-68) ->
// NLopt -> wrap_gradient (:11-17) -> acquisitionfunction(a, model)(x) (src/acquisitionfunctions.jl:4-9) ->
// mean_var = EXT GP.predict_f (src/models/gp.jl:2-5,8) -> functor a(mu, s2) (acquisitionfunctions.jl:24-27,47-50,
// 96,108,111,141).  Per candidate tile (64 columns, one persistent CTA):
//   forward, block row i:  R = k(X_i, x*) - sum_{j<i} L_ij V_j ;  V_i = L_ii^-1 R        (DMMA GEMMs, K = 128 i)
//                          mu += alpha_i' k(X_i, x*),  ssum += colsum(V_i^2)
//   scores:                s2 = max(sf2 - ssum, 0),  a(mu, s2) and its partials, tile arg-max
//   backward (gradient):   R = V_i - sum_{j>i} L_ji' W_j ;  W_i = L_ii^-T R  (same rows of the mirrored factor)
//                          grad_d -= 1/l_d * sum_m (a_mu alpha_m - 2 a_s2 W_m) sf2 psi(r2_m) (z*_d - z_md)
// k(X_i, x*) tiles are generated on the fly and never stored; V/W panels live in a per-CTA scratch that stays
// in L2.  Every candidate's sums are taken in a fixed order that does not depend on its column position, so a
// batched call equals the per-point call bit for bit (reference test/acquisitionfunctions.jl:10).
#include "tma.cuh"
#include "acqfn.cuh"
#include "handle.h"

namespace b200bo {

constexpr int AQ_BM = NB, AQ_BN = TILE_N, AQ_THREADS = 384, AQ_STAGES = 4;
constexpr int RB_STRIDE = NB;   // resident R/G tile [TILE_N][NB], swizzled

struct AcqArgs {
  const double* L; int64_t ld; const double* Linv; const double* LinvT;
  const double* Z; const double* alpha; const double* inv_ell; const double* Xs;
  double* V;
  int64_t M; int N; int nblk; int D; int ntiles;
  double sf2, beta;
  int acq; double p0, p1; unsigned long long seed; int64_t idx_offset;
  double *values, *grad, *mu, *var;
  b200bo_best_t* cta_best;
};

// MODE 0: acquisition step.  MODE 1: the right-hand sides are the columns of I (candidate n of tile t is e_{64 t + n}),
// scores are skipped and the backward pass always runs, leaving Sigma^-1 in the panels (used by the MAP gradient).
//
// Warp roles (384 threads = 3 warpgroups): warps 0-7 are MMA/epilogue consumers (setmaxnreg 240), warp 8 lane 0 is the
// TMA producer (the rest of warpgroup 2 gives its registers away and exits).  Operand chunks (128 or 64 rows x 16 k)
// travel global -> smem by cp.async.bulk.tensor through a AQ_STAGES-deep full/empty mbarrier ring; the consumers never
// execute a CTA-wide barrier inside a contraction.  `vready` orders the consumers' generic-proxy panel writes before the
// producer's async-proxy reads of the same block.
struct AcqMaps { CUtensorMap L, V, Linv, LinvT; };

constexpr int AQ_CONSUMERS = 256;
constexpr uint32_t A_BYTES = AQ_BM * KC * sizeof(double), B_BYTES = AQ_BN * KC * sizeof(double);
constexpr int STAGE_DBL = (AQ_BM + AQ_BN) * KC;
constexpr int CH = NB / KC;   // chunks per 128-wide block

__device__ __forceinline__ void produce_chunk(PipeState& p, double* stages, uint64_t* full, uint64_t* empty, const CUtensorMap* mapA, int ak,
                                              int arow, const CUtensorMap* mapB, int bk, int brow) {
  mbar_wait(&empty[p.s], p.ph ^ 1u);
  double* st = stages + p.s * STAGE_DBL;
  mbar_arrive_expect_tx(&full[p.s], mapB ? A_BYTES + B_BYTES : A_BYTES);
  tma_load_2d(st, mapA, &full[p.s], ak, arow);
  if (mapB) tma_load_2d(st + AQ_BM * KC, mapB, &full[p.s], bk, brow);
  p.next<AQ_STAGES>();
}

template <int FAM, int MODE>
__global__ void __launch_bounds__(384, 1) acq_fused_kernel(const AcqArgs a, const __grid_constant__ AcqMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];              // TMA 128B-swizzle tiles need 1024-byte alignment
  double* stages = reinterpret_cast<double*>(smem_raw);
  const int D = a.D;
  double* Rb = stages + AQ_STAGES * STAGE_DBL;                       // [64][128] swizzled (1024-B aligned)
  double* zx = Rb + TILE_N * RB_STRIDE;                             // [D][128]
  double* zc = zx + D * NB;                                         // [D][64]
  double* alb = zc + D * TILE_N;                                    // [128]
  double* red = alb + NB;                                           // [2][4][64]
  double* c_mu = red + 2 * 4 * TILE_N;                              // [64] each
  double* c_s2 = c_mu + TILE_N;
  double* c_amu = c_s2 + TILE_N;
  double* c_as2 = c_amu + TILE_N;
  double* c_val = c_as2 + TILE_N;
  double* ie = c_val + TILE_N;                                      // [D] inverse length-scales
  uint64_t* full = reinterpret_cast<uint64_t*>(ie + ((D + 1) & ~1));
  uint64_t* empty = full + AQ_STAGES;
  uint64_t* vready = empty + AQ_STAGES;
  double& best_v = *reinterpret_cast<double*>(vready + 1);
  long long& best_i = *reinterpret_cast<long long*>(vready + 2);
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < AQ_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], AQ_CONSUMERS / 32); }
    mbar_init(vready, AQ_CONSUMERS);
    fence_barrier_init();
    best_v = -INFINITY; best_i = -1;
  }
  __syncthreads();
  const bool want_grad = MODE == 1 || a.grad != nullptr;

  if (warp >= AQ_CONSUMERS / 32) {
    // =========================================== TMA producer ===========================================
    reg_dealloc<24>();
    if (warp == AQ_CONSUMERS / 32 && lane == 0) {
      tma_prefetch_desc(&maps.L); tma_prefetch_desc(&maps.V); tma_prefetch_desc(&maps.Linv); tma_prefetch_desc(&maps.LinvT);
      PipeState p;
      uint32_t vph = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int vrow = (MODE == 1 ? tile : (int)blockIdx.x) * TILE_N;
        const int ib0 = MODE == 1 ? (tile * TILE_N) / NB : 0;
        for (int ib = 0; ib < a.nblk; ++ib) {
          for (int jb = ib0; jb < ib - 1; ++jb)
            for (int c = 0; c < CH; ++c) produce_chunk(p, stages, full, empty, &maps.L, jb * NB + c * KC, ib * NB, &maps.V, jb * NB + c * KC, vrow);
          if (ib - 1 >= ib0) {
            mbar_wait(vready, vph); vph ^= 1u;                       // block ib-1 of the panel is written and fenced
            for (int c = 0; c < CH; ++c)
              produce_chunk(p, stages, full, empty, &maps.L, (ib - 1) * NB + c * KC, ib * NB, &maps.V, (ib - 1) * NB + c * KC, vrow);
          }
          if (ib >= ib0)
            for (int c = 0; c < CH; ++c) produce_chunk(p, stages, full, empty, &maps.Linv, c * KC, ib * NB, nullptr, 0, 0);
        }
        if (a.nblk > 0) { mbar_wait(vready, vph); vph ^= 1u; }       // last forward block
        if (want_grad) {
          for (int ib = a.nblk - 1; ib >= 0; --ib) {
            for (int jb = a.nblk - 1; jb > ib + 1; --jb)
              for (int c = 0; c < CH; ++c) produce_chunk(p, stages, full, empty, &maps.L, jb * NB + c * KC, ib * NB, &maps.V, jb * NB + c * KC, vrow);
            if (ib < a.nblk - 1) {
              mbar_wait(vready, vph); vph ^= 1u;                     // W block ib+1
              for (int c = 0; c < CH; ++c)
                produce_chunk(p, stages, full, empty, &maps.L, (ib + 1) * NB + c * KC, ib * NB, &maps.V, (ib + 1) * NB + c * KC, vrow);
            }
            for (int c = 0; c < CH; ++c) produce_chunk(p, stages, full, empty, &maps.LinvT, c * KC, ib * NB, nullptr, 0, 0);
          }
          if (a.nblk > 0) { mbar_wait(vready, vph); vph ^= 1u; }     // W block 0
        }
      }
    }
    return;
  }

  // ============================================= consumers =============================================
  reg_alloc<240>();
  const int wm = warp >> 1, wn = warp & 1;     // 4 x 2 warps, warp tile 32 x 32
  const int g = lane >> 2, q = lane & 3;
  const int rg = rho(g);
  PipeState p;
  const FragAddr fa(rg, q);
  const uint32_t stage0 = smem_u32(stages);
  const uint32_t a_row = (uint32_t)(wm * 32 + rg) * 128u, b_row = (uint32_t)(wn * 32 + rg) * 128u;      // ring tiles: 128-B rows
  const uint32_t rb_row = smem_u32(Rb) + (uint32_t)(wn * 32 + rg) * 1024u;                                  // resident tile: 1024-B rows
  for (int d = tid; d < D; d += AQ_CONSUMERS) ie[d] = a.inv_ell[d];

  // acc[mt][nt][e]  <->  (m, n) = (wm*32 + mt*8 + rg,  wn*32 + nt*8 + q + 4e)
#define ACC_M(mt) (wm * 32 + (mt) * 8 + rg)
#define ACC_N(nt, e) (wn * 32 + (nt) * 8 + q + 4 * (e))

  // acc += sum over `nch` ring chunks of A(128 x 16) * B(64 x 16)^T; B from the ring, or (B_RES) from the resident tile Rb at
  // k = kres0 + 16 c.  Chunks c outside [klo, khi) are released without multiplying (triangular diagonal blocks).
  auto consume = [&](double (&acc)[4][4][2], int nch, bool b_res, int klo, int khi) {
    for (int c = 0; c < nch; ++c) {
      mbar_wait(&full[p.s], p.ph);
      if (c >= klo && c < khi) {
        const uint32_t st = stage0 + (uint32_t)p.s * (STAGE_DBL * 8);
        if (b_res) warp_mma_chunk_t<4, 4, 128, 1024>(acc, st + a_row, rb_row + (uint32_t)c * (KC * 8), fa);
        else warp_mma_chunk_t<4, 4, 128, 128>(acc, st + a_row, st + AQ_BM * KC * 8 + b_row, fa);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[p.s]);
      p.next<AQ_STAGES>();
    }
  };

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    consumer_sync<AQ_CONSUMERS>();
    const int64_t c0 = (int64_t)tile * TILE_N;
    // solve panel [64][ld]: one per resident CTA (MODE 0) or one per tile so the result survives (MODE 1)
    double* Vs = a.V + (int64_t)(MODE == 1 ? tile : (int)blockIdx.x) * TILE_N * a.ld;
    const int ib0 = MODE == 1 ? (int)(c0 / NB) : 0;                 // rows above the identity column are zero
    if (MODE == 0)
      for (int e = tid; e < TILE_N * D; e += AQ_CONSUMERS) {
        const int n = e / D, d = e - n * D;
        int64_t gi = c0 + n;
        if (gi >= a.M) gi = a.M - 1;                                // ragged tail: replicate a valid column
        zc[d * TILE_N + n] = a.Xs[gi * D + d] * ie[d];
      }
    double mu_p[4][2], ss_p[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mu_p[nt][0] = mu_p[nt][1] = ss_p[nt][0] = ss_p[nt][1] = 0.0;

    // ======================================= forward solve =======================================
    for (int ib = 0; ib < a.nblk; ++ib) {
      consumer_sync<AQ_CONSUMERS>();
      if (MODE == 1 && ib < ib0) {
        // never read through TMA (the contractions start at block ib0) and no vready phase: the consumers do not depend
        // on the producer here, so an arrival could run more than one phase ahead of its wait
        for (int e = tid; e < TILE_N * NB; e += AQ_CONSUMERS) Vs[(int64_t)(e >> 7) * a.ld + ib * NB + (e & 127)] = 0.0;
        continue;
      }
      if (MODE == 0) {
        for (int e = tid; e < NB * D; e += AQ_CONSUMERS) {
          const int m = e / D, d = e - m * D;
          zx[d * NB + m] = a.Z[((int64_t)ib * NB + m) * D + d];
        }
        if (tid < NB) alb[tid] = a.alpha[ib * NB + tid];
      }
      consumer_sync<AQ_CONSUMERS>();
      double acc[4][4][2];
      if (MODE == 1) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) acc[mt][nt][e] = ((int64_t)ib * NB + ACC_M(mt) == c0 + ACC_N(nt, e)) ? -1.0 : 0.0;
      } else {
        double r2[4][4][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) r2[mt][nt][0] = r2[mt][nt][1] = 0.0;
        for (int d = 0; d < D; ++d) {
          double xm[4], xn[4][2];
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) xm[mt] = zx[d * NB + ACC_M(mt)];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) { xn[nt][0] = zc[d * TILE_N + ACC_N(nt, 0)]; xn[nt][1] = zc[d * TILE_N + ACC_N(nt, 1)]; }
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const double d0 = xm[mt] - xn[nt][0], d1 = xm[mt] - xn[nt][1];
              r2[mt][nt][0] = fma(d0, d0, r2[mt][nt][0]);
              r2[mt][nt][1] = fma(d1, d1, r2[mt][nt][1]);
            }
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const int m = ACC_M(mt);
          const bool live = ib * NB + m < a.N;
          const double al = alb[m];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const double ks = live ? a.sf2 * kern_phi<FAM>(r2[mt][nt][e]) : 0.0;
              mu_p[nt][e] = fma(al, ks, mu_p[nt][e]);
              acc[mt][nt][e] = -ks;
            }
        }
      }
      consume(acc, (ib - ib0) * CH, false, 0, (ib - ib0) * CH);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            Rb[toff(ACC_N(nt, e), ACC_M(mt), RB_STRIDE)] = -acc[mt][nt][e];
            acc[mt][nt][e] = 0.0;
          }
      consumer_sync<AQ_CONSUMERS>();
      consume(acc, CH, true, 0, 2 * wm + 2);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double v = acc[mt][nt][e];
            Vs[(int64_t)ACC_N(nt, e) * a.ld + ib * NB + ACC_M(mt)] = v;
            ss_p[nt][e] = fma(v, v, ss_p[nt][e]);
          }
      fence_proxy_async_global();
      mbar_arrive(vready);
    }
    // ---- per-candidate reductions: over g (shuffles), then over the 4 warp rows (fixed order) ----
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double m_ = mu_p[nt][e], s_ = ss_p[nt][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          m_ += __shfl_xor_sync(0xffffffffu, m_, o);
          s_ += __shfl_xor_sync(0xffffffffu, s_, o);
        }
        if (g == 0) {
          red[(0 * 4 + wm) * TILE_N + ACC_N(nt, e)] = m_;
          red[(1 * 4 + wm) * TILE_N + ACC_N(nt, e)] = s_;
        }
      }
    consumer_sync<AQ_CONSUMERS>();
    if (MODE == 0 && tid < TILE_N) {
      const int n = tid;
      const int64_t gi = c0 + n;
      const double mu = a.beta + (((red[0 * TILE_N + n] + red[1 * TILE_N + n]) + red[2 * TILE_N + n]) + red[3 * TILE_N + n]);
      const double ss = ((red[4 * TILE_N + n] + red[5 * TILE_N + n]) + red[6 * TILE_N + n]) + red[7 * TILE_N + n];
      const double s2 = fmax(a.sf2 - ss, 0.0);
      double val = mu, amu = 1.0, as2 = 0.0;
      if (a.acq >= 0) {
        const double eps = a.acq == B200BO_ACQ_TS ? philox_normal(a.seed, (unsigned long long)(a.idx_offset + gi)) : 0.0;
        acq_eval(a.acq, a.p0, a.p1, mu, s2, eps, val, amu, as2);
      }
      c_mu[n] = mu; c_s2[n] = s2; c_amu[n] = amu; c_as2[n] = as2; c_val[n] = val;
      if (gi < a.M) {
        if (a.mu) a.mu[gi] = mu;
        if (a.var) a.var[gi] = s2;
        if (a.values) a.values[gi] = val;
      }
    }
    consumer_sync<AQ_CONSUMERS>();
    if (MODE == 0 && tid == 0 && a.acq >= 0) {   // tile arg-max in index order: first strict maximum, NaN never wins
      double bv = best_v; long long bi = best_i;
      for (int n = 0; n < TILE_N; ++n) {
        const int64_t gi = c0 + n;
        if (gi < a.M && better(c_val[n], a.idx_offset + gi, bv, bi)) { bv = c_val[n]; bi = a.idx_offset + gi; }
      }
      best_v = bv; best_i = bi;
    }

    // ======================================= backward solve + gradient =======================================
    if (want_grad) {
      const int gn = tid & (TILE_N - 1), gd = tid >> 6;     // gradient accumulators: candidate gn, dims gd, gd+4, ...
      double gacc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) gacc[k] = 0.0;
      for (int ib = a.nblk - 1; ib >= 0; --ib) {
        consumer_sync<AQ_CONSUMERS>();
        if (MODE == 0) {
          for (int e = tid; e < NB * D; e += AQ_CONSUMERS) {
            const int m = e / D, d = e - m * D;
            zx[d * NB + m] = a.Z[((int64_t)ib * NB + m) * D + d];
          }
          if (tid < NB) alb[tid] = a.alpha[ib * NB + tid];
        }
        double acc[4][4][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) acc[mt][nt][e] = -__ldcg(Vs + (int64_t)ACC_N(nt, e) * a.ld + ib * NB + ACC_M(mt));
        const int nch = (a.nblk - 1 - ib) * CH;
        consume(acc, nch, false, 0, nch);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              Rb[toff(ACC_N(nt, e), ACC_M(mt), RB_STRIDE)] = -acc[mt][nt][e];
              acc[mt][nt][e] = 0.0;
            }
        consumer_sync<AQ_CONSUMERS>();
        consume(acc, CH, true, 2 * wm, CH);
        // W_i -> panel (overwrites V_i), and G = (a_mu alpha - 2 a_s2 W) * sf2 psi(r2) -> Rb
        if (MODE == 1) {
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) Vs[(int64_t)ACC_N(nt, e) * a.ld + ib * NB + ACC_M(mt)] = acc[mt][nt][e];
          fence_proxy_async_global();
          mbar_arrive(vready);
          continue;
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) Vs[(int64_t)ACC_N(nt, e) * a.ld + ib * NB + ACC_M(mt)] = acc[mt][nt][e];
        fence_proxy_async_global();
        mbar_arrive(vready);
        consumer_sync<AQ_CONSUMERS>();           // every warp is done reading Rb (diagonal GEMM) before G overwrites it
        {
          double r2[4][4][2];
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) r2[mt][nt][0] = r2[mt][nt][1] = 0.0;
          for (int d = 0; d < D; ++d) {
            double xm[4], xn[4][2];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) xm[mt] = zx[d * NB + ACC_M(mt)];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { xn[nt][0] = zc[d * TILE_N + ACC_N(nt, 0)]; xn[nt][1] = zc[d * TILE_N + ACC_N(nt, 1)]; }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) {
                const double d0 = xm[mt] - xn[nt][0], d1 = xm[mt] - xn[nt][1];
                r2[mt][nt][0] = fma(d0, d0, r2[mt][nt][0]);
                r2[mt][nt][1] = fma(d1, d1, r2[mt][nt][1]);
              }
          }
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            const int m = ACC_M(mt);
            const bool live = ib * NB + m < a.N;
            const double al = alb[m];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int n = ACC_N(nt, e);
                double phi, psi;
                kern_phi_psi<FAM>(r2[mt][nt][e], phi, psi);
                const double c = fma(c_amu[n], al, -2.0 * c_as2[n] * acc[mt][nt][e]);
                Rb[toff(n, m, RB_STRIDE)] = live ? c * a.sf2 * psi : 0.0;
              }
          }
        }
        consumer_sync<AQ_CONSUMERS>();
        for (int m = 0; m < NB; ++m) {
          const double gv = Rb[toff(gn, m, RB_STRIDE)];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int d = gd + 4 * k;
            if (d < D) gacc[k] = fma(gv, zc[d * TILE_N + gn] - zx[d * NB + m], gacc[k]);
          }
        }
      }
      const int64_t gi = c0 + gn;
      if (MODE == 0 && gi < a.M) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int d = gd + 4 * k;
          if (d < D) a.grad[gi * D + d] = -gacc[k] * ie[d];
        }
      }
    }
  }
#undef ACC_M
#undef ACC_N
  consumer_sync<AQ_CONSUMERS>();
  if (tid == 0) { a.cta_best[blockIdx.x].value = best_v; a.cta_best[blockIdx.x].index = best_i; }
}

// K8 (single-GPU part): reduce the per-CTA bests in index order (value, then lowest index).
__global__ void argmax_reduce_kernel(const b200bo_best_t* __restrict__ cta_best, int n, b200bo_best_t* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double bv = -INFINITY; long long bi = -1;
    for (int i = 0; i < n; ++i) {
      const double v = cta_best[i].value; const long long ix = cta_best[i].index;
      if (ix >= 0 && better(v, ix, bv, bi)) { bv = v; bi = ix; }
    }
    out->value = bv; out->index = bi;
  }
}

size_t acq_smem_bytes(int D) {
  const size_t dbl = (size_t)AQ_STAGES * (AQ_BM + AQ_BN) * KC + (size_t)TILE_N * RB_STRIDE + (size_t)D * NB + (size_t)D * TILE_N + NB +
                     2 * 4 * TILE_N + 5 * TILE_N + ((D + 1) & ~1) + 2 * AQ_STAGES + 4;
  return dbl * sizeof(double);
}

cudaError_t launch_acquire(b200bo_handle_s* h, const AcqLaunch& l) {
  // engine 1 (default): error-free int8-slice GEMM against W = L^-1 on tcgen05 (acq_i8.cu); its int32 accumulators hold 7 N 2^14 < 2^31
  const bool i8 = h->acq_engine < 0 ? acq_i8_default() : h->acq_engine == 1;
  if (i8 && h->Np <= 16384) return launch_acquire_i8(h, l);
  if (l.hXs) {                                                // the DMMA engine has no chunk lanes: one copy in, the outputs out behind the launch
    AcqLaunch d = l;
    d.hXs = nullptr;
    cudaError_t e = l.M > 0 ? cudaMemcpyAsync(const_cast<double*>(l.dXs), l.hXs, sizeof(double) * l.M * h->D, cudaMemcpyHostToDevice, h->stream) : cudaSuccess;
    if (e == cudaSuccess) e = launch_acquire(h, d);
    if (e == cudaSuccess && l.M > 0) {
      if (l.hvalues) e = cudaMemcpyAsync(l.hvalues, l.dvalues, sizeof(double) * l.M, cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess && l.hmu) e = cudaMemcpyAsync(l.hmu, l.dmu, sizeof(double) * l.M, cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess && l.hvar) e = cudaMemcpyAsync(l.hvar, l.dvar, sizeof(double) * l.M, cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess && l.hgrad) e = cudaMemcpyAsync(l.hgrad, l.dgrad, sizeof(double) * l.M * h->D, cudaMemcpyDeviceToHost, h->stream);
    }
    return e;
  }
  AcqArgs a;
  a.L = h->dL; a.ld = h->ld; a.Linv = h->dLinv; a.LinvT = h->dLinvT;
  a.Z = h->dZ; a.alpha = h->dalpha; a.inv_ell = h->dinv_ell; a.Xs = l.dXs; a.V = h->dV;
  a.M = l.M; a.N = (int)h->N; a.nblk = (int)(h->Np / NB); a.D = h->D;
  a.ntiles = (int)((l.M + TILE_N - 1) / TILE_N);
  a.sf2 = exp(2.0 * h->hp.lsigma);
  a.beta = h->mean_kind == B200BO_MEAN_CONST ? h->hp.beta : 0.0;
  a.acq = l.acq_kind; a.p0 = l.p0; a.p1 = l.p1; a.seed = l.seed; a.idx_offset = l.idx_offset;
  a.values = l.dvalues; a.grad = l.dgrad; a.mu = l.dmu; a.var = l.dvar;
  a.cta_best = h->dcta_best;
  if (a.ntiles == 0) return cudaSuccess;
  const int grid = (int)(a.ntiles < h->nslots ? a.ntiles : h->nslots);
  const size_t smem = acq_smem_bytes(h->D);
  AcqMaps maps;
  maps.L = h->tmL; maps.V = h->tmV; maps.Linv = h->tmLinv; maps.LinvT = h->tmLinvT;
#define B200BO_ACQ(F)                                                                                   \
  do {                                                                                                  \
    cudaFuncSetAttribute(acq_fused_kernel<F, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    acq_fused_kernel<F, 0><<<grid, AQ_THREADS, smem, h->stream>>>(a, maps);                                  \
  } while (0)
  switch (h->fam) {
    case FAM_SE: B200BO_ACQ(FAM_SE); break;
    case FAM_MAT12: B200BO_ACQ(FAM_MAT12); break;
    case FAM_MAT32: B200BO_ACQ(FAM_MAT32); break;
    default: B200BO_ACQ(FAM_MAT52); break;
  }
#undef B200BO_ACQ
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (l.dbest && l.acq_kind >= 0) {
    argmax_reduce_kernel<<<1, 32, 0, h->stream>>>(h->dcta_best, grid, l.dbest);
    h->launches++;
    e = cudaGetLastError();
  }
  return e;
}

// Sigma^-1 = L^-T L^-1 by the same fused forward/backward sweep applied to the columns of I.  Result: h->dV viewed
// as [Np][ld] (row j = column j of the symmetric inverse).
cudaError_t launch_kinv_solve(b200bo_handle_s* h) {
  AcqArgs a = {};
  a.L = h->dL; a.ld = h->ld; a.Linv = h->dLinv; a.LinvT = h->dLinvT;
  a.V = h->dV; a.M = h->Np; a.N = (int)h->N; a.nblk = (int)(h->Np / NB); a.D = 0;
  a.ntiles = (int)(h->Np / TILE_N);
  a.acq = -1; a.cta_best = h->dcta_best;
  if (a.ntiles == 0) return cudaSuccess;
  if (a.ntiles > h->nslots) return cudaErrorInvalidValue;
  const size_t smem = acq_smem_bytes(0);
  cudaFuncSetAttribute(acq_fused_kernel<FAM_SE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  AcqMaps maps;
  maps.L = h->tmL; maps.V = h->tmV; maps.Linv = h->tmLinv; maps.LinvT = h->tmLinvT;
  acq_fused_kernel<FAM_SE, 1><<<a.ntiles, AQ_THREADS, smem, h->stream>>>(a, maps);
  h->launches++;
  return cudaGetLastError();
}

}  // namespace b200bo
