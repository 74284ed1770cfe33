// kmat.cu -- K1: kernel-matrix assembly  Sigma = K(X,X) + (e^{2 logNoise} + eps) I   (FP64, HBM-write bound).
//
// Replaces EXT GaussianProcesses.jl `update_cK!` (cov(kernel, x, x) + noise on the diagonal) reached from
// update!(model, x, y) (reference src/models/gp.jl:11-18).  Math: SURVEY.md Appendix A "Parametrisation"/"Fit".
//
// One CTA per 64x64 tile of the LOWER triangle of tiles; the tile is computed once and stored twice (tile and its transpose, the
// thread's 4x4 block transposed in registers) so the output is exactly symmetric and every global store is one full 32-byte
// sector per lane (STG.E.256).  Algorithmic bytes: 8 N^2 written + 8 N D read.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "tma.cuh"
#include "handle.h"

namespace b200bo {

constexpr int KT = 64;   // tile edge

// K1 staging block of one aligned group of 64 points, KS = ceil(D/4) k-steps of the FP64 MMA (doubles):
//   [B fragments 256 KS][|z|^2/2  64][A fragments 256 KS][|z|^2/2  64]
// A fragments (the group as tile ROWS):  [wr 2][ks][g 8][q 4][mt 4] = z[32 wr + 4 g + mt][4 ks + q]
// B fragments (the group as tile COLS):  [wc 4][ks][g 8][q 4][nt 2] = z[16 wc + 4 (g>>1) + 2 nt + (g&1)][4 ks + q]
// so lane (g, q) of warp (wr, wc) fetches its DMMA operands for one k-step with three conflict-free LDS.128 and ends up owning the
// 4 x 4 block rows 32 wr + 4 g .., cols 16 wc + 4 q .. of the tile.  Dimensions D..4 KS-1 stay zero (memset at allocation).
__host__ __device__ __forceinline__ int64_t kblock_doubles(int D) { return (int64_t)(8 * ((D + 3) / 4) + 2) * KT; }

// Z = X / ell, point-major (the TMA-fed kernels) and as K1 staging blocks.  One thread per point.
__global__ void scale_inputs_kernel(const double* __restrict__ X, const double* __restrict__ inv_ell, double* __restrict__ Z,
                                    double* __restrict__ Zk, int D, int64_t n0, int64_t n1) {
  const int64_t p = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n1) return;
  const int KS = (D + 3) / 4;
  double* blk = Zk + (p >> 6) * kblock_doubles(D);
  double* Bf = blk;
  double* Af = blk + 256 * KS + KT;
  const int c = (int)(p & 63);
  const int wr = c >> 5, ga = (c & 31) >> 2, mt = c & 3;
  const int wc = c >> 4, cc = c & 15, gb = 2 * (cc >> 2) + (cc & 1), nt = (cc >> 1) & 1;
  double h = 0.0;
  for (int d = 0; d < D; ++d) {
    const double x = X[p * D + d];
    const double z = (x - inv_ell[D + d]) * inv_ell[d];     // K1 only: centred on the mid-range of the data (Gram-form rounding ~ eps |z|^2)
    const int ks = d >> 2, q = d & 3;
    Z[p * D + d] = x * inv_ell[d];
    Af[(((wr * KS + ks) * 8 + ga) * 4 + q) * 4 + mt] = z;
    Bf[(((wc * KS + ks) * 8 + gb) * 4 + q) * 2 + nt] = z;
    h = fma(z, z, h);
  }
  blk[256 * KS + c] = 0.5 * h;
  blk[512 * KS + KT + c] = 0.5 * h;
}

// 32-byte vector store (sm_100: STG.E.256): one full sector per lane
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
template <bool WIDE>
__device__ __forceinline__ void st_row4(double* p, double a, double b, double c, double d) {
  if (WIDE) { st256(p, a, b, c, d); return; }
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
  *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
}

// 1-D TMA bulk copy global -> shared, completing on an mbarrier (SASS UBLKCP)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// sf2 * phi for x = -r^2/2 (<= 0): T[j] = sf2 * 2^(j/16) in shared memory
template <int FAM>
__device__ __forceinline__ double kern_from_neg_half_r2(double x, const double* T) {
  if (FAM == FAM_SE) return exp_neg16(x, T);
  const double r = sqrt(-2.0 * x);
  if (FAM == FAM_MAT12) return exp_neg16(-r, T);
  if (FAM == FAM_MAT32) { const double s = 1.7320508075688772 * r; return (1.0 + s) * exp_neg16(-s, T); }
  const double s = 2.23606797749979 * r;
  return (1.0 + s + s * s * (1.0 / 3.0)) * exp_neg16(-s, T);
}

__device__ __forceinline__ void tile_of(int t, int& bi, int& bj) {   // linear lower-triangle tile index -> (bi >= bj)
  bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  while (bi * (bi + 1) / 2 > t) --bi;
  bj = t - bi * (bi + 1) / 2;
}

// Persistent CTAs over the 64x64 tiles of the lower triangle of tiles.  The two staging blocks of the NEXT tile are prefetched by
// 1-D TMA bulk copies (SASS UBLKCP) into the other half of a double buffer (mbarrier complete_tx) while the current tile is
// computed, so no warp ever waits on a global load.  -r^2/2 = z_i.z_j - |z_i|^2/2 - |z_j|^2/2 with the dot products on the FP64
// tensor pipe (DMMA.8x8x4, 16 per warp and tile at D = 8), the exponential by a 16-entry table + degree-7 polynomial (12 FP64
// instructions); the tile is stored twice (itself and its register-transposed mirror), every store instruction of a warp writing
// eight full 128-byte lines (STG.E.256).
template <int FAM, bool WIDE>
__global__ void __launch_bounds__(256, 4) kmat_kernel(const double* __restrict__ Zk, int ntiles, int N, int Np, int D, double sf2,
                                                      double post, double noise, int pad_identity, double* __restrict__ K, int64_t ld) {
  extern __shared__ __align__(128) double sm[];
  const int KS = (D + 3) / 4;
  const int part = 256 * KS + KT;                // doubles per operand: fragments + |z|^2/2
  const int64_t gblk = 2 * (int64_t)part;        // doubles per staging block in global memory
  double* T = sm;                                // [16]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16);   // [0..1] full (TMA bytes landed), [2..3] empty (all 8 warps done reading)
  double* buf = sm + 32;                         // [2][2][part]: {A of the row group, B of the col group} x double buffer
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int wr = w >> 2, wc = w & 3, g = lane >> 2, q = lane & 3;
  const int ty = 8 * wr + g, tx = 4 * wc + q;    // the thread owns rows 4 ty .. 4 ty + 3, cols 4 tx .. 4 tx + 3 of the tile
  if (tid < 16) T[tid] = sf2 * exp2((double)tid * 0.0625);
  if (tid == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 8); mbar_init(&bar[3], 8);
    fence_barrier_init();
  }
  __syncthreads();
  const uint32_t part_bytes = (uint32_t)part * 8u;
  int t = blockIdx.x;
  if (tid == 0 && t < ntiles) {
    int bi, bj; tile_of(t, bi, bj);
    mbar_arrive_expect_tx(&bar[0], 2 * part_bytes);
    bulk_load(buf, Zk + bi * gblk + part, part_bytes, &bar[0]);
    bulk_load(buf + part, Zk + bj * gblk, part_bytes, &bar[0]);
  }
  const int lim = pad_identity ? Np : N;
  for (int it = 0; t < ntiles; t += gridDim.x, ++it) {
    const int b = it & 1;
    int bi, bj; tile_of(t, bi, bj);
    if (tid == 0 && t + (int)gridDim.x < ntiles) {     // prefetch the next tile once every warp has released its buffer
      if (it > 0) mbar_wait(&bar[2 + (b ^ 1)], (uint32_t)((it - 1) >> 1) & 1u);
      int ni, nj; tile_of(t + gridDim.x, ni, nj);
      double* nb = buf + (b ^ 1) * 2 * part;
      mbar_arrive_expect_tx(&bar[b ^ 1], 2 * part_bytes);
      bulk_load(nb, Zk + ni * gblk + part, part_bytes, &bar[b ^ 1]);
      bulk_load(nb + part, Zk + nj * gblk, part_bytes, &bar[b ^ 1]);
    }
    const double* za = buf + b * 2 * part;
    const double* zb = za + part;
    mbar_wait(&bar[b], (uint32_t)(it >> 1) & 1u);
    double acc[4][2][2];            // [mt][nt][e] = element (row 4 ty + mt, col 4 tx + 2 nt + e)
    double hsum_hi;                 // largest |z_i|^2/2 + |z_j|^2/2 of the thread's 4 x 4 block: scale of the Gram-form rounding error
    {
      const double2 ha0 = *reinterpret_cast<const double2*>(za + 256 * KS + 4 * ty), ha1 = *reinterpret_cast<const double2*>(za + 256 * KS + 4 * ty + 2);
      const double2 hb0 = *reinterpret_cast<const double2*>(zb + 256 * KS + 4 * tx), hb1 = *reinterpret_cast<const double2*>(zb + 256 * KS + 4 * tx + 2);
      const double a4[4] = {ha0.x, ha0.y, ha1.x, ha1.y}, b4[4] = {hb0.x, hb0.y, hb1.x, hb1.y};
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[mt][c >> 1][c & 1] = -(a4[mt] + b4[c]);
      hsum_hi = fmax(fmax(a4[0], a4[1]), fmax(a4[2], a4[3])) + fmax(fmax(b4[0], b4[1]), fmax(b4[2], b4[3]));
    }
    {
      const double* Ap = za + ((wr * KS * 8 + g) * 4 + q) * 4;
      const double* Bp = zb + ((wc * KS * 8 + g) * 4 + q) * 2;
#pragma unroll 2
      for (int ks = 0; ks < KS; ++ks) {
        const double2 a01 = *reinterpret_cast<const double2*>(Ap + ks * 128), a23 = *reinterpret_cast<const double2*>(Ap + ks * 128 + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(Bp + ks * 64);
        const double af[4] = {a01.x, a01.y, a23.x, a23.y}, bf[2] = {b01.x, b01.y};
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
      }
    }
    if (FAM != FAM_SE) {   // exp(-r^2/2) passes a residual of 1e-13 through unchanged; the Matern families take its square root
      // (near-)coincident points: z_i.z_j - |z_i|^2/2 - |z_j|^2/2 has cancelled >= 24 bits and its residual is rounding noise of either
      // sign (exact duplicates are appended by `repetitions` and re-proposed by the search).  Recompute those few elements from the
      // coordinate differences, as K6 and the elastic append do: r = 0 exactly for duplicates, full precision next to them.
      double amax = acc[0][0][0];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int c = 0; c < 4; ++c) amax = fmax(amax, acc[mt][c >> 1][c & 1]);
      if (__any_sync(0xffffffffu, amax > -0x1p-24 * hsum_hi)) {
        const double* Ar = za + ((wr * KS * 8 + g) * 4) * 4;
        const double* Bc = zb + (wc * KS * 8 * 4) * 2;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (!(acc[mt][c >> 1][c & 1] > -0x1p-24 * (za[256 * KS + 4 * ty + mt] + zb[256 * KS + 4 * tx + c]))) continue;
            const int gb = 2 * q + (c & 1), nt = c >> 1;
            double r2 = 0.0;
            for (int d = 0; d < 4 * KS; ++d) {
              const int ks = d >> 2, qq = d & 3;
              const double df = Ar[(ks * 32 + qq) * 4 + mt] - Bc[((ks * 8 + gb) * 4 + qq) * 2 + nt];
              r2 = fma(df, df, r2);
            }
            acc[mt][c >> 1][c & 1] = -0.5 * r2;
          }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar[2 + b]);   // this warp is done with buffer b (no CTA-wide barrier in the loop)
    if (bi == bj && ty == tx) {     // a point against itself: r = 0 exactly
#pragma unroll
      for (int a = 0; a < 4; ++a) acc[a][a >> 1][a & 1] = 0.0;
    }
    double v[4][4];
    if ((bi + 1) * KT <= N) {
      // ---- interior tile (the common case): no bounds logic at all ----
      double* p = K + (int64_t)(bi * KT + 4 * ty) * ld + bj * KT + 4 * tx;
      const bool dg = bi == bj && ty == tx;
#pragma unroll
      for (int a = 0; a < 4; ++a) {   // each row leaves as soon as it is computed: stores spread over the FP64 phase
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) v[a][bb] = kern_from_neg_half_r2<FAM>(fmin(acc[a][bb >> 1][bb & 1], 0.0), T);
        if (post != 1.0) {            // extreme signal variance: not folded into the table
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) v[a][bb] *= post;
        }
        if (dg) v[a][a] += noise;
        st_row4<WIDE>(p + (int64_t)a * ld, v[a][0], v[a][1], v[a][2], v[a][3]);
      }
      if (bi != bj) {   // mirrored tile: the thread's 4 x 4 block transposed in registers
        double* m = K + (int64_t)(bj * KT + 4 * tx) * ld + bi * KT + 4 * ty;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) st_row4<WIDE>(m + (int64_t)bb * ld, v[0][bb], v[1][bb], v[2][bb], v[3][bb]);
      }
      continue;
    }
    // ---- edge tile: ragged N and the identity padding ----
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int gi = bi * KT + 4 * ty + a;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int gj = bj * KT + 4 * tx + bb;
        double val;
        if (gi < N && gj < N) {
          val = post * kern_from_neg_half_r2<FAM>(fmin(acc[a][bb >> 1][bb & 1], 0.0), T);
          if (gi == gj) val += noise;
        } else {
          val = (pad_identity && gi == gj) ? 1.0 : 0.0;
        }
        v[a][bb] = val;
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int gi = bi * KT + 4 * ty + a, gj = bj * KT + 4 * tx;
      if (gi < lim) {
        double* p = K + (int64_t)gi * ld + gj;
        if (gj + 3 < lim) {
          st_row4<WIDE>(p, v[a][0], v[a][1], v[a][2], v[a][3]);
        } else {
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) if (gj + bb < lim) p[bb] = v[a][bb];
        }
      }
    }
    if (bi != bj) {
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int gi = bj * KT + 4 * tx + bb, gj = bi * KT + 4 * ty;
        if (gi < lim) {
          double* p = K + (int64_t)gi * ld + gj;
          if (gj + 3 < lim) {
            st_row4<WIDE>(p, v[0][bb], v[1][bb], v[2][bb], v[3][bb]);
          } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) if (gj + a < lim) p[a] = v[a][bb];
          }
        }
      }
    }
  }
}

cudaError_t launch_scale_inputs(b200bo_handle_s* h, int64_t n0, int64_t n1) {
  if (n1 <= n0) return cudaSuccess;
  const int blocks = (int)((n1 - n0 + 127) / 128);
  scale_inputs_kernel<<<blocks, 128, 0, h->stream>>>(h->dX, h->dinv_ell, h->dZ, h->dZk, h->D, n0, n1);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_kmat(b200bo_handle_s* h, double* dK, int64_t ld, int64_t N, int64_t Np, double noise, bool pad_identity) {
  const int64_t lim = pad_identity ? Np : N;
  const int T = (int)((lim + KT - 1) / KT);
  const int ntiles = T * (T + 1) / 2;
  if (ntiles == 0) return cudaSuccess;
  const size_t smem = (size_t)(32 + 2 * kblock_doubles(h->D)) * sizeof(double);
  const int wide = (((uintptr_t)dK & 31) == 0 && (ld & 3) == 0) ? 1 : 0;
  double sf2 = exp(2.0 * h->hp.lsigma), post = 1.0;
  if (!(sf2 > 1e-12 && sf2 < 1e12)) { post = sf2; sf2 = 1.0; }
  // persistent grid: every SM holds the same number of CTAs and CTA c takes tiles c, c + grid, ... (tile counts per SM differ by <= 1)
  static const int knob = getenv("B200BO_KMAT_CTAS") ? atoi(getenv("B200BO_KMAT_CTAS")) : 4;   // developer knob
  int per_sm = 1, grid = 1;
#define B200BO_KMAT2(F, W)                                                                                   \
  do {                                                                                                       \
    cudaFuncSetAttribute(kmat_kernel<F, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kmat_kernel<F, W>, 256, smem);                    \
    grid = std::min(ntiles, std::max(1, std::min(per_sm, knob)) * h->num_sms);                               \
    kmat_kernel<F, W><<<grid, 256, smem, h->stream>>>(h->dZk, ntiles, (int)N, (int)Np, h->D, sf2, post, noise, pad_identity ? 1 : 0, dK, ld); \
  } while (0)
#define B200BO_KMAT(F)                   \
  do {                                   \
    if (wide) B200BO_KMAT2(F, true);     \
    else B200BO_KMAT2(F, false);         \
  } while (0)
  switch (h->fam) {
    case FAM_SE: B200BO_KMAT(FAM_SE); break;
    case FAM_MAT12: B200BO_KMAT(FAM_MAT12); break;
    case FAM_MAT32: B200BO_KMAT(FAM_MAT32); break;
    default: B200BO_KMAT(FAM_MAT52); break;
  }
#undef B200BO_KMAT
#undef B200BO_KMAT2
  h->launches++;
  return cudaGetLastError();
}

}  // namespace b200bo
