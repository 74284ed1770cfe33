// acq_i8.cu -- K6 on the 5th-generation tensor cores: the batched acquisition step as an error-free int8-slice GEMM (tcgen05.mma
// kind::i8, TMEM accumulators) against the explicit inverse factor.
//
// Replaces, per BO iteration, the reference's multi-restart search: acquire_max (src/acquisition.jl:54-68) -> NLopt -> wrap_gradient
// (:11-17) -> acquisitionfunction(a, model)(x) (src/acquisitionfunctions.jl:4-9) -> mean_var = EXT GP.predict_f (src/models/gp.jl:2-5,8)
// -> functor a(mu, s2) (acquisitionfunctions.jl:24-27,47-50,96,108,111,141).
//
// tcgen05 has no FP64 kind, and a triangular SOLVE per candidate tile chains one block row behind the other.  Both go away with
//     v = U^-T k*  =  W k*,   W = L^-1  (lower triangular, built once per factor by recursive block inversion, kinv.cu)
// which is ONE dependency-free GEMM  V[M x N] = K*[M x N] W^T  over all candidates.  FP64 accuracy on an int8 tensor pipe: every
// operand row is split once into 7 balanced radix-256 digits with a power-of-two scale (W: per row, at fit time; k*: fixed scale
// 2^ceil(log2 sf2), inside the kernel-evaluation pass), and the product is the 28 EXACT slice products p + q <= 6 accumulated in
// int32 by anti-diagonal (Ozaki scheme; truncation 2^-51 max|W_i| sf2 per term).  sigma^2 = sf2 - |v|^2 is then a sum of squares, with no
// cancellation inside; measured against a long-double solve: |d sigma^2| <= 3e-14 at N = 2048 (oracle/ozaki.py, tests/test_oracle.py).
//
// Per chunk of CH candidates (CH N 7 bytes of k* slices stay L2-resident), three launches on one stream:
//   kstar_slice_kernel   k(X_j, x*) for a 128 x 64 block: r^2 by coordinate differences, phi, mu partial = alpha_j' k*, digits -> slices
//   acq_i8_gemm_kernel   persistent CTAs over (128-candidate tile, 64-row tile of W): TMA (SWIZZLE_128B, 3-D maps over the slices)
//                        -> 448 x nkb UMMAs 128 x 64 x 32 -> 7 TMEM accumulators -> exact int64 recombination -> per-candidate
//                        partial sums of v^2 (MODE 0) or w = Sigma^-1 k* itself (MODE 1, the gradient's second product)
//   acq_finish_kernel    mu, sigma^2 in a fixed summation order, functor + partials, outputs, per-block arg-max (first strict maximum)
// and, with a gradient, acq_i8_gemm_kernel<1> against the slices of Sigma^-1 + acq_grad_kernel (closed form of SURVEY App. A).
// Every per-candidate quantity is computed in an order that does not depend on the candidate's position or on its neighbours (the
// integer products are exact), so a batched call equals the per-point call bit for bit (reference test/acquisitionfunctions.jl:10).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "umma.cuh"
#include "umma_issue.cuh"
#include "acqfn.cuh"
#include "handle.h"

namespace b200bo {

constexpr int A8_S = 7;                  // slices per FP64 value
constexpr int A8_BM = 128;               // candidates per tile  = UMMA M = TMEM lanes
constexpr int A8_BN = 64;                // rows of W / Sigma^-1 per tile = UMMA N
constexpr uint32_t A8_A_BYTES = A8_BM * 128, A8_B_BYTES = A8_BN * 128;
constexpr size_t A8_SMEM = 2 * A8_S * A8_B_BYTES + A8_S * A8_A_BYTES + 1024;      // 224 KB of operand stages + barriers
constexpr int A8_THREADS = 320;          // producer warp, MMA warp, 8 epilogue warps
constexpr int64_t A8_CHUNK_BYTES = 64ll << 20;   // k* slices of one chunk; two chunks are in flight (L2 holds 126 MB)

// byte k of 16 digit words -> one 16-byte vector (element u in byte u)
__device__ __forceinline__ uint4 gather_byte(const unsigned long long (&dg)[16], int k) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = k < 4 ? (uint32_t)dg[4 * i + u] : (uint32_t)(dg[4 * i + u] >> 32);
    const uint32_t sel = (uint32_t)(k & 3) | ((uint32_t)(4 + (k & 3)) << 4);     // [x.k, y.k, -, -]
    const uint32_t t0 = __byte_perm(x[0], x[1], sel), t1 = __byte_perm(x[2], x[3], sel);
    w[i] = __byte_perm(t0, t1, 0x5410);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- FP64 rows -> 7 int8 slices + per-row scale 2^(e-6).  One warp per row; tri = 1: only columns [0, 128 (row / 128 + 1)) exist (W = L^-1),
//      tri = 2: only columns [128 (row / 128), ncols_full) exist (W^T); the other columns are never read by the products. ----
__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ A, int64_t lda, int nrows, int ncols_full, int tri,
                                                         uint8_t* __restrict__ Sl, int64_t pitch, int64_t slice_stride,
                                                         double* __restrict__ Se) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const int ncols = tri == 1 ? (((row >> 7) + 1) << 7) : ncols_full;       // a multiple of 128
  const int clo = tri == 2 ? ((row >> 7) << 7) : 0;
  const double* src = A + (int64_t)row * lda;
  double m = 0.0;
  for (int c = clo + 2 * lane; c < ncols; c += 64) {
    const double2 v = *reinterpret_cast<const double2*>(src + c);
    m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int e = (m > 0.0 && m < INFINITY) ? ilogb(m) + 1 : 0;             // |x| < 2^e
  const double sc = ldexp(1.0, 54 - e);                                   // |x sc| <= 2^54, an integer wherever x has the row's top exponent
  for (int c0 = clo + 16 * lane; c0 < ncols; c0 += 512) {
    unsigned long long dg[16];
#pragma unroll
    for (int u = 0; u < 16; u += 2) {
      const double2 v = *reinterpret_cast<const double2*>(src + c0 + u);
      dg[u] = i8_digits(__double2ll_rn(v.x * sc));
      dg[u + 1] = i8_digits(__double2ll_rn(v.y * sc));
    }
#pragma unroll
    for (int s = 0; s < A8_S; ++s)
      *reinterpret_cast<uint4*>(Sl + (int64_t)s * slice_stride + (int64_t)row * pitch + c0) = gather_byte(dg, 6 - s);
  }
  if (lane == 0) Se[row] = ldexp(1.0, e - 6);
}

// ---- k* = k(X, x*) for one 128 (observations) x 64 (candidates) block: values -> digits -> slices, and the partial posterior mean ----
struct KsArgs {
  const double* Z; const double* alpha; const double* inv_ell; const double* Xs;
  uint8_t* Bs; double* MuP;
  double* G;                    // gradient launches: sf2 psi(r2) [Np][CH] (psi = -2 dphi/dr2), so that the gradient pass evaluates no kernel
  int64_t M, c0, CH, Kp;        // candidates in the call, first candidate of the chunk, chunk capacity (rows per slice), bytes per slice row
  int N, D;
  double sf2, qscale;           // qscale = 2^54 / S, S = the fixed power-of-two scale of k*
};

template <int FAM, int DP, bool WG>      // WG: also write sf2 psi (gradient launches); a template flag keeps the value path at 70 registers
__global__ void __launch_bounds__(256) kstar_slice_kernel(const KsArgs a) {
  extern __shared__ __align__(16) uint8_t ks_smem[];
  double* zx = reinterpret_cast<double*>(ks_smem);            // [128][DP]
  double* alb = zx + 128 * DP;                                // [128]
  double* red = alb + 128;                                    // [8][64]
  uint8_t* stage = reinterpret_cast<uint8_t*>(red + 8 * 64);  // [7][64 candidates][128 B], 16-byte chunks XOR-swizzled by (candidate & 7)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int jb = blockIdx.x, D = a.D;
  const int64_t t0 = (int64_t)blockIdx.y * 64;                // first candidate of the block inside the chunk
  for (int e = tid; e < 128 * DP; e += 256) {
    const int m = e / DP, d = e - m * DP;
    zx[e] = d < D ? a.Z[((int64_t)jb * 128 + m) * D + d] : 0.0;
  }
  if (tid < 128) alb[tid] = a.alpha[jb * 128 + tid];
  __syncthreads();
  const int j0 = 16 * w;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int n = lane + 32 * pass;
    int64_t gi = a.c0 + t0 + n;
    if (gi >= a.M) gi = a.M - 1;                              // ragged tail: replicate a valid column (never written back)
    double zc[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) zc[d] = d < D ? a.Xs[gi * D + d] * a.inv_ell[d] : 0.0;
    unsigned long long dg[16];
    double mu = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const double* zr = zx + (j0 + j) * DP;
      double r2 = 0.0;
#pragma unroll
      for (int d = 0; d < DP; d += 2) {
        const double2 x = *reinterpret_cast<const double2*>(zr + d);
        const double u0 = x.x - zc[d], u1 = x.y - zc[d + 1];
        r2 = fma(u0, u0, r2);
        r2 = fma(u1, u1, r2);
      }
      const bool live = jb * 128 + j0 + j < a.N;
      double phi, psi = 0.0;
      if (WG) kern_phi_psi<FAM>(r2, phi, psi); else phi = kern_phi<FAM>(r2);       // the same phi either way, bit for bit
      const double ks = live ? a.sf2 * phi : 0.0;
      if (WG) __stcs(a.G + ((int64_t)jb * 128 + j0 + j) * a.CH + t0 + n, live ? a.sf2 * psi : 0.0);
      mu = fma(alb[j0 + j], ks, mu);
      dg[j] = i8_digits(__double2ll_rn(ks * a.qscale));
    }
    red[w * 64 + n] = mu;
#pragma unroll
    for (int s = 0; s < A8_S; ++s)
      *reinterpret_cast<uint4*>(stage + ((s * 64 + n) << 7) + ((w ^ (n & 7)) << 4)) = gather_byte(dg, 6 - s);
  }
  __syncthreads();
  if (tid < 64) {
    double mu = red[tid];
#pragma unroll
    for (int k = 1; k < 8; ++k) mu += red[k * 64 + tid];
    a.MuP[(int64_t)jb * a.CH + t0 + tid] = mu;
  }
  for (int e = tid; e < A8_S * 64 * 8; e += 256) {            // 8 consecutive threads write one 128-byte row segment
    const int c = e & 7, n = (e >> 3) & 63, s = e >> 9;
    const uint4 v = *reinterpret_cast<const uint4*>(stage + ((s * 64 + n) << 7) + ((c ^ (n & 7)) << 4));
    *reinterpret_cast<uint4*>(a.Bs + ((int64_t)s * a.CH + t0 + n) * a.Kp + (int64_t)jb * 128 + 16 * c) = v;
  }
}

// ---- the GEMM ----
struct A8Maps { CUtensorMap A, B; };   // A: k* slices [7][CH][Kp] in 128-row boxes, B: W or Sigma^-1 slices [7][cap][cap] in 64-row boxes

// work items sorted by decreasing length (it descending), dealt to the CTAs in snake order so that every CTA gets the same mix
template <int MODE>
__device__ __forceinline__ bool a8_item(int r, int G, int b, int total, int nct, int nit, int& it, int& ct) {
  const int pos = r * G + ((r & 1) ? G - 1 - b : b);
  if (pos >= total) return false;
  it = MODE == 2 ? pos / nct : nit - 1 - pos / nct;          // MODE 2 items get longer towards it = 0
  ct = pos - (pos / nct) * nct;
  return true;
}

// MODE 0: out = SsP[2 nit][CH], partial sums over 32 rows of v^2, v = W k*; k-blocks 0 .. it/2 (W is lower triangular).  With `vs` the
//         epilogue also slices v (fixed scale 2^ev >= 2 sigma_f) into vs[7][CH][Kp]: the A operand of the gradient's second product
// MODE 1: out = WgT[Np][CH], w = Sigma^-1 k* from the slices of Sigma^-1; all k-blocks (kept for reference; MODE 2 does half the work)
// MODE 2: out = WgT[Np][CH], w = W^T v from the slices of v and of W^T; k-blocks it/2 .. nkb_full-1 (W^T is upper triangular)
// TS: the k* slices reach the MMAs through tensor memory (tcgen05.cp into two 32-column buffers behind the accumulators) instead of
// being re-read from shared memory by each of the up to 7 MMAs per k-step that use them: the shared-memory port (128 B/clk), which
// bounds the SS form at 840 KB per k-block, carries 504 KB.
template <int MODE, bool TS>
__global__ void __launch_bounds__(A8_THREADS, 1) acq_i8_gemm_kernel(const double* __restrict__ Be, double sBk, int nct, int nit, int nkb_full,
                                                                     int64_t CH, double* __restrict__ out, uint8_t* __restrict__ vs, int64_t Kp,
                                                                     double vq, const __grid_constant__ A8Maps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;                                     // [2][7 slices][64 rows x 128 B]: W / Sigma^-1, double buffered per k-block
  uint8_t* sA = smem_raw + 2 * A8_S * A8_B_BYTES;             // [7 slices][128 rows x 128 B]: k*, stage p = slice p of the current k-block
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + A8_S * A8_A_BYTES);
  uint64_t *afull = bars, *aempty = bars + 7, *bfull = bars + 14, *bempty = bars + 16, *tfull = bars + 18, *tempty = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int G = gridDim.x, b = blockIdx.x, total = nct * nit;
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  if (tid == 0) {
    for (int s = 0; s < 3; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 8);                                     // one arrival per epilogue warp
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {                                        // ===== TMA producer =====
      tma_prefetch_desc(&maps.A); tma_prefetch_desc(&maps.B);
      uint32_t kcnt = 0;                                      // k-blocks so far: B buffer kcnt & 1, phase of every A stage kcnt & 1
      int it, ct;
      for (int r = 0; a8_item<MODE>(r, G, b, total, nct, nit, it, ct); ++r) {
        const int arow = ct * A8_BM, brow = it * A8_BN;
        const int kb0 = MODE == 2 ? it / 2 : 0, nkb = MODE == 0 ? it / 2 + 1 : nkb_full;
        for (int kb = kb0; kb < nkb; ++kb, ++kcnt) {
          const int bs = kcnt & 1;
          mbar_wait_or_trap(&bempty[bs], ((kcnt >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bfull[bs], A8_S * A8_B_BYTES);
          for (int q = 0; q < A8_S; ++q) tma_load_3d(sB + (bs * A8_S + q) * A8_B_BYTES, &maps.B, &bfull[bs], kb * 128, brow, q);
          for (int g = 0; g < 3; ++g) {                         // slices {0,1}, {2,3}, {4,5,6}: one full/empty barrier pair per group
            const int p0 = 2 * g, p1 = g == 2 ? A8_S : 2 * g + 2;
            mbar_wait_or_trap(&aempty[g], (kcnt & 1u) ^ 1u);
            mbar_arrive_expect_tx(&afull[g], (uint32_t)(p1 - p0) * A8_A_BYTES);
            for (int p = p0; p < p1; ++p) tma_load_3d(sA + p * A8_A_BYTES, &maps.A, &afull[g], kb * 128, arow, p);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                        // ===== MMA issuer =====
      // D = s32, A = B = signed 8-bit, both K-major, N = 64, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(A8_BN >> 3) << 17) | ((uint32_t)(A8_BM >> 4) << 24);
      const uint64_t dA0 = umma_desc_sw128(smem_u32(sA)), dB0 = umma_desc_sw128(smem_u32(sB));
      uint32_t kcnt = 0;
      int it, ct;
#ifdef A8_PROF
      long long w_t = 0, w_a = 0, w_b = 0, w_a0 = 0, w_ap[7] = {0,0,0,0,0,0,0}, t_all = clock64(), tq; int nitems = 0;
#define A8_TIMED(acc_, stmt) do { tq = clock64(); stmt; acc_ += clock64() - tq; } while (0)
#define A8_EXTRA(P) { const long long dt_ = clock64() - tq; w_ap[P] += dt_; if (kb == 0) w_a0 += dt_; }
#else
#define A8_TIMED(acc_, stmt) do { stmt; } while (0)
#define A8_EXTRA(P)
#endif
      for (int r = 0; a8_item<MODE>(r, G, b, total, nct, nit, it, ct); ++r) {
        const int kb0 = MODE == 2 ? it / 2 : 0, nkb = MODE == 0 ? it / 2 + 1 : nkb_full;
#ifdef A8_PROF
        ++nitems;
#endif
        A8_TIMED(w_t, mbar_wait_or_trap(tempty, ((uint32_t)r & 1u) ^ 1u));   // the epilogue has read the previous item's accumulators
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int kb = kb0; kb < nkb; ++kb, ++kcnt) {
          const uint32_t bs = kcnt & 1u, aph = kcnt & 1u, acc = kb > kb0;
          const uint64_t db = dB0 + (uint64_t)(bs * ((A8_S * A8_B_BYTES) >> 4));
          A8_TIMED(w_b, mbar_wait_or_trap(&bfull[bs], (kcnt >> 1) & 1u));
          // slice p of k*: one asm block issues its 4 (7 - p) MMAs (umma_issue.cuh); tcgen05.commit frees the stage behind them
#define A8_WAIT(G)                                                                                                \
          A8_TIMED(w_a, mbar_wait_or_trap(&afull[G], aph)); A8_EXTRA(G)                                           \
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#define A8_ISSUE(P)                                                                                               \
          if (TS) umma_i8_issue_ts<P, A8_BN, (A8_B_BYTES >> 4), 1>(tmem, tmem + 448u + 32u * ((P) & 1), dA0 + (uint64_t)((P) * (A8_A_BYTES >> 4)), db, idesc, (P) ? 1u : acc); \
          else umma_i8_issue<P, A8_BN, (A8_B_BYTES >> 4), 1>(tmem, dA0 + (uint64_t)((P) * (A8_A_BYTES >> 4)), db, idesc, (P) ? 1u : acc);

          A8_WAIT(0) A8_ISSUE(0) A8_ISSUE(1) umma_commit(&aempty[0]);
          A8_WAIT(1) A8_ISSUE(2) A8_ISSUE(3) umma_commit(&aempty[1]);
          A8_WAIT(2) A8_ISSUE(4) A8_ISSUE(5) A8_ISSUE(6) umma_commit(&aempty[2]);
#undef A8_ISSUE
#undef A8_WAIT
          umma_commit(&bempty[bs]);
        }
        umma_commit(tfull);                                   // every MMA of the item retired: accumulators complete
      }
#ifdef A8_PROF
      if (b == 0 && kcnt > 0)
        printf("acq_i8_gemm<%d> CTA0: %u k-blocks %d items, %lld cycles per k-block; waits per k-block: TMEM %lld, A %lld (of which first k-block of an item %lld; by slice %lld %lld %lld %lld %lld %lld %lld), B %lld\n", MODE,
               kcnt, nitems, (clock64() - t_all) / kcnt, w_t / kcnt, w_a / kcnt, w_a0 / kcnt, w_ap[0] / kcnt, w_ap[1] / kcnt, w_ap[2] / kcnt, w_ap[3] / kcnt, w_ap[4] / kcnt, w_ap[5] / kcnt, w_ap[6] / kcnt, w_b / kcnt);
#endif
    }
  } else {
    // ===== epilogue: 8 warps; warp w reads TMEM lanes 32 (w & 3) .. +31 (= candidates of the tile), columns 32 half .. +31 (= rows of W) =====
    const int g4 = warp & 3, half = (warp - 2) >> 2;
    const int m = 32 * g4 + lane;
    int it, ct;
    for (int r = 0; a8_item<MODE>(r, G, b, total, nct, nit, it, ct); ++r) {
      const int brow = it * A8_BN + 32 * half;
      mbar_wait_or_trap(tfull, (uint32_t)r & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // sum_d a_d 2^(-8d) in two exact 64-bit integer groups (|a_d| < 2^31):  (((a0 256 + a1) 256 + a2) 256 + a3) 2^-24 + ((a4 256 + a5) 256 + a6) 2^-48
      double acc[32];
#pragma unroll 1
      for (int grp = 1; grp >= 0; --grp) {                    // the small group first
        long long H[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) H[c] = 0;
        const int nd = grp ? A8_S - 4 : 4;
#pragma unroll 1
        for (int dd = 0; dd < nd; ++dd) {
          uint32_t rr[32];
          tmem_ld32(tmem + ((uint32_t)(32 * g4) << 16) + (uint32_t)((4 * grp + dd) * A8_BN + 32 * half), rr);
#pragma unroll
          for (int c = 0; c < 32; ++c) H[c] = (H[c] << 8) + (long long)(int32_t)rr[c];
        }
        if (grp == 0) {                                       // TMEM has been read: the next item's MMAs may overwrite it
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
        }
        const double wgt = grp ? 0x1p-48 : 0x1p-24;
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = grp ? (double)H[c] * wgt : fma((double)H[c], wgt, acc[c]);
      }
      if (MODE == 0) {
        double ss = 0.0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          acc[c] *= sBk * __ldg(Be + brow + c);
          ss = fma(acc[c], acc[c], ss);
        }
        out[(int64_t)(2 * it + half) * CH + (int64_t)ct * A8_BM + m] = ss;
        if (vs) {                                             // v -> 7 int8 slices, 32 consecutive rows of W = 32 bytes per slice
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            unsigned long long dg[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) dg[u] = i8_digits(__double2ll_rn(acc[16 * hh + u] * vq));
#pragma unroll
            for (int s = 0; s < A8_S; ++s)
              *reinterpret_cast<uint4*>(vs + ((int64_t)s * CH + (int64_t)ct * A8_BM + m) * Kp + brow + 16 * hh) = gather_byte(dg, 6 - s);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) __stcs(out + (int64_t)(brow + c) * CH + (int64_t)ct * A8_BM + m, acc[c] * (sBk * __ldg(Be + brow + c)));   // read once, by the gradient pass
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

// ---- scores: one thread per candidate of the chunk ----
struct FinArgs {
  const double* MuP; const double* SsP;
  int64_t CH, M, c0, idx_offset;
  int nblk, nss;                // partials per candidate: nblk of the mean, nss of |v|^2
  double sf2, beta;
  int acq; double p0, p1; unsigned long long seed;
  double *values, *mu, *var, *amu, *as2;
  b200bo_best_t* cta_best;      // one per block of the launch
};

__device__ __forceinline__ void best_merge(double& bv, long long& bi, double v, long long i) {
  if (i >= 0 && (bi < 0 || v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
}

// 256 threads = 32 candidates x 8 partial-sum classes: class g adds the partials k = g, g + 8, ... in ascending order, the eight class
// sums are then added in ascending g (a fixed tree, the same for every candidate wherever it sits in the batch)
__global__ void __launch_bounds__(256) acq_finish_kernel(const FinArgs f) {
  __shared__ double smu[8][32], sss[8][32];
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  const int64_t n = (int64_t)blockIdx.x * 32 + lane, gi = f.c0 + n;
  double musum = 0.0, ss = 0.0;
  for (int k = g; k < f.nblk; k += 8) musum += f.MuP[(int64_t)k * f.CH + n];
  for (int k = g; k < f.nss; k += 8) ss += f.SsP[(int64_t)k * f.CH + n];
  smu[g][lane] = musum; sss[g][lane] = ss;
  __syncthreads();
  if (g != 0) return;
  musum = smu[0][lane]; ss = sss[0][lane];
#pragma unroll
  for (int k = 1; k < 8; ++k) { musum += smu[k][lane]; ss += sss[k][lane]; }
  const double mu = f.beta + musum;
  const double s2 = fmax(f.sf2 - ss, 0.0);
  double val = mu, amu = 1.0, as2 = 0.0;
  if (f.acq >= 0) {
    const double eps = f.acq == B200BO_ACQ_TS ? philox_normal(f.seed, (unsigned long long)(f.idx_offset + gi)) : 0.0;
    acq_eval(f.acq, f.p0, f.p1, mu, s2, eps, val, amu, as2);
  }
  const bool valid = gi < f.M;
  if (f.amu) { f.amu[n] = amu; f.as2[n] = as2; }
  if (valid) {
    if (f.mu) f.mu[gi] = mu;
    if (f.var) f.var[gi] = s2;
    if (f.values) f.values[gi] = val;
  }
  // first strict maximum in index order: the largest value, then the lowest index; NaN and -Inf never win (acquire_max: `f > maxf`)
  double bv = -INFINITY; long long bi = -1;
  if (valid && f.acq >= 0 && val > -INFINITY) { bv = val; bi = f.idx_offset + gi; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    best_merge(bv, bi, ov, oi);
  }
  if (lane == 0) { f.cta_best[blockIdx.x].value = bv; f.cta_best[blockIdx.x].index = bi; }
}

__global__ void __launch_bounds__(256) argmax_blocks_kernel(const b200bo_best_t* __restrict__ in, int n, b200bo_best_t* __restrict__ out) {
  __shared__ double sv[8];
  __shared__ long long si[8];
  double bv = -INFINITY; long long bi = -1;
  for (int i = threadIdx.x; i < n; i += 256) best_merge(bv, bi, in[i].value, in[i].index);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    best_merge(bv, bi, ov, oi);
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) best_merge(bv, bi, sv[k], si[k]);
    out->value = bv; out->index = bi;
  }
}

// ---- gradient: grad_d = -1/l_d sum_m (a_mu alpha_m - 2 a_s2 w_m) sf2 psi(r2_m) (z*_d - z_md)    ----
// w = Sigma^-1 k* comes from the second slice product (WgT [Np][CH]), sf2 psi from the kernel-evaluation pass (G [Np][CH]): this pass
// is D + 3 FMAs per (observation, candidate) pair.  Block = 32 candidates (lanes: coalesced rows of WgT / G) x 4 groups of observations;
// blockIdx.y takes a contiguous range of 128-blocks and leaves a partial sum, which grad_reduce_kernel adds in ascending order.
struct GradArgs {
  const double* Z; const double* alpha; const double* inv_ell; const double* Xs; const double* WgT; const double* G; const double* amu; const double* as2;
  double* part;                 // [splits][CH][D]
  double* grad;
  int64_t M, c0, CH;
  int N, D, nblk, nsplit;
};

template <int DP>
__global__ void __launch_bounds__(128) acq_grad_kernel(const GradArgs a) {
  extern __shared__ __align__(16) uint8_t gr_smem[];
  double* zx = reinterpret_cast<double*>(gr_smem);            // [128][DP]
  double* alb = zx + 128 * DP;                                // [128]
  double* red = alb + 128;                                    // [4][32][DP + 1]
  const int tid = threadIdx.x, n = tid & 31, mg = tid >> 5, D = a.D;
  const int64_t cn = (int64_t)blockIdx.x * 32 + n;
  int64_t gi = a.c0 + cn;
  if (gi >= a.M) gi = a.M - 1;
  double zc[DP], gacc[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) { zc[d] = d < D ? a.Xs[gi * D + d] * a.inv_ell[d] : 0.0; gacc[d] = 0.0; }
  const double amu = a.amu[cn], as2m2 = -2.0 * a.as2[cn];
  const int per = (a.nblk + a.nsplit - 1) / a.nsplit;
  const int jb0 = blockIdx.y * per, jb1 = min(a.nblk, jb0 + per);
  for (int jb = jb0; jb < jb1; ++jb) {
    __syncthreads();
    for (int e = tid; e < 128 * DP; e += 128) {
      const int m = e / DP, d = e - m * DP;
      zx[e] = d < D ? a.Z[((int64_t)jb * 128 + m) * D + d] : 0.0;
    }
    alb[tid] = a.alpha[jb * 128 + tid];
    __syncthreads();
    const double* wp = a.WgT + ((int64_t)jb * 128 + 32 * mg) * a.CH + cn;
    const double* gp = a.G + ((int64_t)jb * 128 + 32 * mg) * a.CH + cn;
#pragma unroll 2
    for (int j = 0; j < 32; ++j) {
      const int m = 32 * mg + j;
      const double c = fma(amu, alb[m], as2m2 * __ldcs(wp + (int64_t)j * a.CH)) * __ldcs(gp + (int64_t)j * a.CH);   // G is zero in the padding rows
      const double* zr = zx + m * DP;
#pragma unroll
      for (int d = 0; d < DP; d += 2) {
        const double2 x = *reinterpret_cast<const double2*>(zr + d);
        gacc[d] = fma(c, zc[d] - x.x, gacc[d]);
        gacc[d + 1] = fma(c, zc[d + 1] - x.y, gacc[d + 1]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int d = 0; d < DP; ++d) red[(mg * 32 + n) * (DP + 1) + d] = gacc[d];
  __syncthreads();
  for (int e = tid; e < 32 * D; e += 128) {
    const int nn = e / D, d = e - nn * D;
    const double s = ((red[nn * (DP + 1) + d] + red[(32 + nn) * (DP + 1) + d]) + red[(64 + nn) * (DP + 1) + d]) + red[(96 + nn) * (DP + 1) + d];
    a.part[((int64_t)blockIdx.y * a.CH + (int64_t)blockIdx.x * 32 + nn) * D + d] = s;
  }
}

__global__ void __launch_bounds__(256) grad_reduce_kernel(const GradArgs a, int64_t mc) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= mc * a.D) return;
  const int64_t cn = e / a.D; const int d = (int)(e - cn * a.D);
  double s = 0.0;
  for (int k = 0; k < a.nsplit; ++k) s += a.part[((int64_t)k * a.CH + cn) * a.D + d];
  a.grad[(a.c0 + cn) * a.D + d] = -s * a.inv_ell[d];
}

// ------------------------------------------------------------------------------------------------------------------------------------
cudaError_t make_map3d_u8(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1);

bool acq_i8_default() {
  static const bool on = !(getenv("B200BO_ACQ_I8") && atoi(getenv("B200BO_ACQ_I8")) == 0);   // B200BO_ACQ_I8=0 selects the DMMA kernel
  return on;
}

static int64_t chunk_of(int64_t Np, int64_t knob_mb) {
  const int64_t budget = knob_mb > 0 ? (knob_mb << 20) : A8_CHUNK_BYTES;
  int64_t ch = budget / (A8_S * Np) / 128 * 128;
  return std::max<int64_t>(128, std::min<int64_t>(ch, 8192));
}

// W = L^-1 (and, for the gradient, Sigma^-1) of the current factor as int8 slices.  Once per factor.
static cudaError_t ensure_slices(b200bo_handle_s* h, bool want_grad) {
  const int64_t cap = h->cap;
  const int Np = (int)h->Np;
  cudaError_t e = cudaSuccess;
  if (!(h->acq_ready & 1)) {
    if (!h->dWs) {
      e = cudaMalloc(&h->dWs, (size_t)A8_S * cap * cap);
      if (e == cudaSuccess) e = cudaMalloc(&h->dWe, sizeof(double) * cap);
      if (e == cudaSuccess) e = make_map3d_u8(&h->tmWsB, h->dWs, (uint64_t)cap, (uint64_t)cap, A8_S, 128, A8_BN);
      if (e != cudaSuccess) { cudaFree(h->dWs); cudaFree(h->dWe); h->dWs = nullptr; h->dWe = nullptr; return e; }
    }
    e = launch_linv(h);                                       // W row-major in h->dKi, W^T in h->dWT
    if (e != cudaSuccess) return e;
    h->wt_valid = true;
    slice_rows_kernel<<<(Np + 7) / 8, 256, 0, h->stream>>>(h->dKi, h->ld, Np, Np, 1, reinterpret_cast<uint8_t*>(h->dWs), cap, cap * cap, h->dWe);
    h->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    h->acq_ready = 1;
  }
  if (want_grad && !(h->acq_ready & 2)) {                      // the gradient's second product w = W^T v needs the rows of W^T
    if (!h->dKs) {
      e = cudaMalloc(&h->dKs, (size_t)A8_S * cap * cap);
      if (e == cudaSuccess) e = cudaMalloc(&h->dKe, sizeof(double) * cap);
      if (e == cudaSuccess) e = make_map3d_u8(&h->tmKsB, h->dKs, (uint64_t)cap, (uint64_t)cap, A8_S, 128, A8_BN);
      if (e != cudaSuccess) { cudaFree(h->dKs); cudaFree(h->dKe); h->dKs = nullptr; h->dKe = nullptr; return e; }
    }
    if (!h->wt_valid) {                                       // W^T was overwritten (a MAP sweep ran on these buffers): rebuild it
      if ((e = launch_linv(h)) != cudaSuccess) return e;
      h->wt_valid = true;
    }
    slice_rows_kernel<<<(Np + 7) / 8, 256, 0, h->stream>>>(h->dWT, h->ld, Np, Np, 2, reinterpret_cast<uint8_t*>(h->dKs), cap, cap * cap, h->dKe);
    h->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    h->acq_ready |= 2;
  }
  return cudaSuccess;
}

// chunk buffers, TWO sets (consecutive chunks alternate between two stream lanes): k* slices, partial sums, (gradient) w / sf2 psi /
// partial gradients; sized for the current Np
static cudaError_t ensure_chunk_buffers(b200bo_handle_s* h, int64_t CH, bool want_grad, int64_t nblocks) {
  const int64_t Np = std::max<int64_t>(h->Np, NB);
  cudaError_t e = cudaSuccess;
  const size_t lane_bs = (size_t)A8_S * CH * Np;              // per lane: k* slices, and behind them (gradient) the slices of v
  const size_t need_bs = (want_grad ? 4 : 2) * lane_bs;
  if (need_bs > h->bs_bytes) {
    cudaFree(h->dBs); h->dBs = nullptr; h->bs_bytes = 0; h->bs_np = 0;
    if ((e = cudaMalloc(&h->dBs, need_bs)) != cudaSuccess) return e;
    h->bs_bytes = need_bs;
  }
  if (h->bs_np != Np || h->bs_ch != CH || (want_grad && !h->bs_grad)) {
    const int nmaps = h->bs_bytes >= 4 * lane_bs ? 4 : 2;
    for (int ln = 0; ln < nmaps; ++ln)
      if ((e = make_map3d_u8(&h->tmBsA[ln], static_cast<uint8_t*>(h->dBs) + ln * lane_bs, (uint64_t)Np, (uint64_t)CH, A8_S, 128, A8_BM)) != cudaSuccess) return e;
    h->bs_np = Np; h->bs_ch = CH; h->bs_grad = nmaps == 4;
  }
  const size_t need_p = 2 * sizeof(double) * (size_t)CH * (size_t)(Np / NB + Np / 32 + 2);
  if (need_p > h->part_bytes) {
    cudaFree(h->dMuP); h->dMuP = nullptr; h->part_bytes = 0;
    if ((e = cudaMalloc(&h->dMuP, need_p)) != cudaSuccess) return e;
    h->part_bytes = need_p;
  }
  if (want_grad) {
    const size_t need_w = 2 * sizeof(double) * (size_t)CH * (size_t)(2 * Np + 16 * h->D);      // w = Sigma^-1 k*, sf2 psi, <= 16 partial gradients
    if (need_w > h->wg_bytes) {
      cudaFree(h->dWg); h->dWg = nullptr; h->wg_bytes = 0;
      if ((e = cudaMalloc(&h->dWg, need_w)) != cudaSuccess) return e;
      h->wg_bytes = need_w;
    }
  }
  if (nblocks > h->nbest2) {
    cudaFree(h->dcta_best2); h->dcta_best2 = nullptr; h->nbest2 = 0;
    if ((e = cudaMalloc(&h->dcta_best2, sizeof(b200bo_best_t) * (size_t)nblocks)) != cudaSuccess) return e;
    h->nbest2 = nblocks;
  }
  for (auto& ev : h->acq_ev)
    if (!ev && (e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
  return cudaSuccess;
}

template <int FAM>
static cudaError_t launch_kstar(b200bo_handle_s* h, cudaStream_t st, const KsArgs& a, int nblk, int ntile64) {
  const int D = h->D;
  const int DP = D <= 4 ? 4 : D <= 8 ? 8 : D <= 16 ? 16 : 32;
  const size_t smem = sizeof(double) * (128 * DP + 128 + 8 * 64) + A8_S * 64 * 128;
  const dim3 grid(nblk, ntile64);
#define B200BO_KS(DPV)                                                                                                     \
  do {                                                                                                                     \
    if (a.G) {                                                                                                             \
      cudaFuncSetAttribute(kstar_slice_kernel<FAM, DPV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
      kstar_slice_kernel<FAM, DPV, true><<<grid, 256, smem, st>>>(a);                                                      \
    } else {                                                                                                               \
      cudaFuncSetAttribute(kstar_slice_kernel<FAM, DPV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
      kstar_slice_kernel<FAM, DPV, false><<<grid, 256, smem, st>>>(a);                                                     \
    }                                                                                                                      \
  } while (0)
  if (DP == 4) B200BO_KS(4); else if (DP == 8) B200BO_KS(8); else if (DP == 16) B200BO_KS(16); else B200BO_KS(32);
#undef B200BO_KS
  h->launches++;
  return cudaGetLastError();
}

static cudaError_t launch_grad(b200bo_handle_s* h, cudaStream_t st, const GradArgs& a, int nblocks, int64_t mc) {
  const int D = h->D;
  const int DP = D <= 4 ? 4 : D <= 8 ? 8 : D <= 16 ? 16 : 32;
  const size_t smem = sizeof(double) * (128 * DP + 128 + 4 * 32 * (DP + 1));
  const dim3 grid(nblocks, a.nsplit);
#define B200BO_GR(DPV)                                                                                         \
  do {                                                                                                         \
    cudaFuncSetAttribute(acq_grad_kernel<DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    acq_grad_kernel<DPV><<<grid, 128, smem, st>>>(a);                                                   \
  } while (0)
  if (DP == 4) B200BO_GR(4); else if (DP == 8) B200BO_GR(8); else if (DP == 16) B200BO_GR(16); else B200BO_GR(32);
#undef B200BO_GR
  grad_reduce_kernel<<<(unsigned)((mc * D + 255) / 256), 256, 0, st>>>(a, mc);
  h->launches += 2;
  return cudaGetLastError();
}

// bench instrumentation (knob "acq_gemm_timing"): an event before and after every slice-product launch
static void gemm_event(b200bo_handle_s* h, cudaStream_t st) {
  if (!h->acq_time_gemm) return;
  if (h->gemm_ev_used == (int)h->gemm_ev.size()) { cudaEvent_t ev; if (cudaEventCreate(&ev) != cudaSuccess) return; h->gemm_ev.push_back(ev); }
  cudaEventRecord(h->gemm_ev[h->gemm_ev_used++], st);
}

// The acquisition step over l.M candidate columns on the tcgen05 path.  Same contract as launch_acquire (acq.cu).
cudaError_t launch_acquire_i8(b200bo_handle_s* h, const AcqLaunch& l) {
  const int64_t M = l.M;
  if (M == 0) return cudaSuccess;
  const int64_t Np = h->Np, Npb = std::max<int64_t>(Np, NB);
  const int nblk = (int)(Np / NB), nit = (int)(Np / A8_BN);
  const bool want_grad = l.dgrad != nullptr;
  cudaError_t e = cudaSuccess;
  if (nblk > 0 && (e = ensure_slices(h, want_grad)) != cudaSuccess) return e;
  // chunks of equal size, at most the L2 budget each; consecutive chunks run on two stream lanes with their own buffers, so the
  // kernel-evaluation pass and the scores of one chunk fill the SMs that the slice product of the other leaves idle at its tail
  const int64_t CHmax = chunk_of(Npb, h->acq_chunk_mb);
  int64_t nchunks = (M + CHmax - 1) / CHmax;
  const int64_t CH = std::min<int64_t>(CHmax, ((M + nchunks - 1) / nchunks + 127) / 128 * 128);
  nchunks = (M + CH - 1) / CH;
  const int64_t nblocks_total = nchunks * (CH / 32);
  if ((e = ensure_chunk_buffers(h, CH, want_grad, nblocks_total)) != cudaSuccess) return e;
  const size_t lane_bs = (size_t)A8_S * CH * Npb, lane_p = (size_t)CH * (size_t)(Npb / NB + Npb / 32 + 2), lane_w = (size_t)CH * (size_t)(2 * Npb + 16 * h->D);
  const double sf2 = exp(2.0 * h->hp.lsigma);
  int e2 = 0;
  frexp(sf2, &e2);                                            // sf2 = f 2^e2, f in [0.5, 1): the fixed scale of k* is S = 2^e2 > sf2
  const double qscale = ldexp(1.0, 54 - e2), sBk = ldexp(1.0, e2 - 6);
  int ev = 0;
  frexp(sqrt(sf2), &ev);                                      // |v_i| <= sigma_f < 2^ev: the slices of v use the fixed scale 2^(ev+1)
  const double vq = ldexp(1.0, 54 - (ev + 1)), sVk = ldexp(1.0, ev + 1 - 6);
  static const bool ts = getenv("B200BO_ACQ_TS") && atoi(getenv("B200BO_ACQ_TS")) == 1;   // developer knob: 1 = A through tensor memory (measured slower: the tcgen05.cp copies cost more than the A-collector reuse saves)
  const bool one_lane = h->acq_lanes == 1 || h->acq_time_gemm;
  h->gemm_ev_used = 0;
  cudaFuncSetAttribute(acq_i8_gemm_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A8_SMEM);
  cudaFuncSetAttribute(acq_i8_gemm_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A8_SMEM);
  cudaFuncSetAttribute(acq_i8_gemm_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A8_SMEM);
  cudaFuncSetAttribute(acq_i8_gemm_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A8_SMEM);
  const bool two = nchunks > 1 && !one_lane && h->stream2 != nullptr;
  cudaStream_t lanes[2] = {h->stream, two ? h->stream2 : h->stream};
  if (two) {                                                  // lane 1 starts behind everything already queued on the handle's stream
    if ((e = cudaEventRecord(h->acq_ev[0], h->stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(h->stream2, h->acq_ev[0], 0)) != cudaSuccess) return e;
  }
  int64_t blk0 = 0;
  int64_t ci = 0;
  for (int64_t c0 = 0; c0 < M; c0 += CH, ++ci) {
    const int ln = two ? (int)(ci & 1) : 0;
    cudaStream_t st = lanes[ln];
    uint8_t* dBs = static_cast<uint8_t*>(h->dBs) + ln * lane_bs;
    uint8_t* dVs = want_grad ? static_cast<uint8_t*>(h->dBs) + (2 + ln) * lane_bs : nullptr;
    double* dMuP = h->dMuP + ln * lane_p;
    double* dSsP = dMuP + (size_t)CH * (size_t)(Npb / NB);
    double* dAmu = dSsP + (size_t)CH * (size_t)(Npb / 32);
    double* dAs2 = dAmu + CH;
    double* dWg = want_grad ? h->dWg + ln * lane_w : nullptr;
    A8Maps mapsW, mapsK;
    mapsW.A = h->tmBsA[ln]; mapsW.B = h->tmWsB;
    mapsK.A = h->tmBsA[want_grad ? 2 + ln : ln]; mapsK.B = h->tmKsB;        // A = slices of v, B = slices of W^T
    const int64_t mc = std::min<int64_t>(CH, M - c0);
    const int64_t mcp = (mc + 127) / 128 * 128;
    const int nct = (int)(mcp / A8_BM);
    if (l.hXs && (e = cudaMemcpyAsync(const_cast<double*>(l.dXs) + c0 * h->D, l.hXs + c0 * h->D, sizeof(double) * mc * h->D, cudaMemcpyHostToDevice, st)) != cudaSuccess)
      return e;                                               // this chunk's candidates arrive on its own lane, under the other lane's kernels
    if (nblk > 0) {
      KsArgs k;
      k.Z = h->dZ; k.alpha = h->dalpha; k.inv_ell = h->dinv_ell; k.Xs = l.dXs; k.Bs = dBs; k.MuP = dMuP;
      k.G = want_grad ? dWg + (size_t)CH * (size_t)Np : nullptr;
      k.M = M; k.c0 = c0; k.CH = CH; k.Kp = Np; k.N = (int)h->N; k.D = h->D; k.sf2 = sf2; k.qscale = qscale;
      switch (h->fam) {
        case FAM_SE: e = launch_kstar<FAM_SE>(h, st, k, nblk, (int)(mcp / 64)); break;
        case FAM_MAT12: e = launch_kstar<FAM_MAT12>(h, st, k, nblk, (int)(mcp / 64)); break;
        case FAM_MAT32: e = launch_kstar<FAM_MAT32>(h, st, k, nblk, (int)(mcp / 64)); break;
        default: e = launch_kstar<FAM_MAT52>(h, st, k, nblk, (int)(mcp / 64)); break;
      }
      if (e != cudaSuccess) return e;
      const int total = nct * nit;
      const int grid = std::min(total, h->num_sms);
      gemm_event(h, st);
      if (ts) acq_i8_gemm_kernel<0, true><<<grid, A8_THREADS, A8_SMEM, st>>>(h->dWe, sBk, nct, nit, nblk, CH, dSsP, dVs, Np, vq, mapsW);
      else acq_i8_gemm_kernel<0, false><<<grid, A8_THREADS, A8_SMEM, st>>>(h->dWe, sBk, nct, nit, nblk, CH, dSsP, dVs, Np, vq, mapsW);
      gemm_event(h, st);
      h->launches++;
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    FinArgs f;
    f.MuP = dMuP; f.SsP = dSsP; f.CH = CH; f.M = M; f.c0 = c0; f.idx_offset = l.idx_offset;
    f.nblk = nblk; f.nss = 2 * nit; f.sf2 = sf2; f.beta = h->mean_kind == B200BO_MEAN_CONST ? h->hp.beta : 0.0;
    f.acq = l.acq_kind; f.p0 = l.p0; f.p1 = l.p1; f.seed = l.seed;
    f.values = l.dvalues; f.mu = l.dmu; f.var = l.dvar; f.amu = want_grad ? dAmu : nullptr; f.as2 = want_grad ? dAs2 : nullptr;
    f.cta_best = h->dcta_best2 + blk0;
    acq_finish_kernel<<<(unsigned)(mcp / 32), 256, 0, st>>>(f);
    h->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    blk0 += mcp / 32;
    if (l.hXs) {                                              // the chunk's outputs go home behind its own kernels
      if (l.hvalues && (e = cudaMemcpyAsync(l.hvalues + c0, l.dvalues + c0, sizeof(double) * mc, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
      if (l.hmu && (e = cudaMemcpyAsync(l.hmu + c0, l.dmu + c0, sizeof(double) * mc, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
      if (l.hvar && (e = cudaMemcpyAsync(l.hvar + c0, l.dvar + c0, sizeof(double) * mc, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    }
    if (want_grad) {
      if (nblk == 0) {
        if ((e = cudaMemsetAsync(l.dgrad + c0 * h->D, 0, sizeof(double) * mc * h->D, st)) != cudaSuccess) return e;
        if (l.hgrad && (e = cudaMemcpyAsync(l.hgrad + c0 * h->D, l.dgrad + c0 * h->D, sizeof(double) * mc * h->D, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        continue;
      }
      const int total = nct * nit;
      const int grid = std::min(total, h->num_sms);
      gemm_event(h, st);
      if (ts) acq_i8_gemm_kernel<2, true><<<grid, A8_THREADS, A8_SMEM, st>>>(h->dKe, sVk, nct, nit, nblk, CH, dWg, nullptr, Np, 0.0, mapsK);
      else acq_i8_gemm_kernel<2, false><<<grid, A8_THREADS, A8_SMEM, st>>>(h->dKe, sVk, nct, nit, nblk, CH, dWg, nullptr, Np, 0.0, mapsK);
      gemm_event(h, st);
      h->launches++;
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      GradArgs g;
      g.Z = h->dZ; g.alpha = h->dalpha; g.inv_ell = h->dinv_ell; g.Xs = l.dXs; g.WgT = dWg; g.G = dWg + (size_t)CH * (size_t)Np;
      g.amu = dAmu; g.as2 = dAs2; g.part = dWg + 2 * (size_t)CH * (size_t)Np; g.grad = l.dgrad;
      g.M = M; g.c0 = c0; g.CH = CH; g.N = (int)h->N; g.D = h->D; g.nblk = nblk;
      const int nb = (int)((mc + 31) / 32);
      // a function of N only (the summation order of a candidate's gradient must not depend on the batch); two blocks of observations per
      // split up to 8 splits: the lock-step L-BFGS rounds of a BO loop are 16 candidates against N in the hundreds, where this kernel was 2 CTAs
      g.nsplit = std::max(1, std::min(8, (nblk + 1) / 2));
      if ((e = launch_grad(h, st, g, nb, mc)) != cudaSuccess) return e;
      if (l.hgrad && (e = cudaMemcpyAsync(l.hgrad + c0 * h->D, l.dgrad + c0 * h->D, sizeof(double) * mc * h->D, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    }
  }
  if (two) {                                                  // join: the handle's stream continues behind lane 1
    if ((e = cudaEventRecord(h->acq_ev[1], h->stream2)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(h->stream, h->acq_ev[1], 0)) != cudaSuccess) return e;
  }
  if (l.dbest && l.acq_kind >= 0) {
    argmax_blocks_kernel<<<1, 256, 0, h->stream>>>(h->dcta_best2, (int)blk0, l.dbest);
    h->launches++;
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace b200bo
