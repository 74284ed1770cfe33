// syrk_i8.cu -- K4 on the 5th-generation tensor cores: the K = 512 trailing update  A_ij -= L_i,P L_j,P^T  of the blocked Cholesky
// (chol.cu) as an error-free int8-slice product (Ozaki scheme) on tcgen05.mma kind::i8 with TMEM accumulators.
//
// tcgen05 has no FP64 kind, so each FP64 operand row is split ONCE per outer panel into 7 int8 slices with a per-row power-of-two
// scale:   x_k = 2^(e-6) sum_p d_p[k] 2^(-8p),  d_0 in [-64, 64], d_p in [-128, 127] (p >= 1, balanced signed digits): 6 + 6 x 8 = 54
// bits, the splitting is exact (slice_panel_kernel).  Then  x.y = 2^(ex+ey-12) sum_{p+q<=6} 2^(-8(p+q)) <d_p, d'_q>  + O(2^-51 max|x| max|y|) per term:
// 28 exact int8 dot products, accumulated in int32 by anti-diagonal d = p + q (|sum| <= 7 * 512 * 2^14 < 2^26) in 7 TMEM
// accumulators of 64 columns.
// Persistent CTAs (one per SM) over the 128 x 64 tiles of the trailing matrix (the same tile set as the DMMA kernel), 320 threads:
//   warp 0   TMA producer: per 128-byte k-block the 7 slices of the 64 B-rows (56 KB, double buffered) and, through a 4-deep
//            ring, the 7 slices of the 128 A-rows one at a time (16 KB each), all SWIZZLE_128B tiles by cp.async.bulk.tensor.3d;
//   warp 1   allocates the 512 TMEM columns and issues the 28 x 4 x 4 = 448 MMAs (128 x 64 x 32) of the tile; tcgen05.commit
//            releases the shared-memory stages and finally publishes the accumulators;
//   warps 2-9 read the accumulators back (tcgen05.ld 32x32b), recombine the anti-diagonals exactly in two 64-bit integer groups, convert
//            once per group, apply the row scales and subtract from A (only the lower triangle of diagonal tiles is touched).
// Every mbarrier wait is bounded (trap instead of hanging the GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "umma.cuh"
#include "umma_issue.cuh"
#include "handle.h"

namespace b200bo {

constexpr int I8_S = 7;                 // slices per FP64 value
constexpr int I8_K = 512;               // bytes of K per slice row = columns of one outer panel
constexpr int I8_BM = 128, I8_BN = 64;
constexpr int I8_ASTAGES = 4;
constexpr uint32_t I8_A_BYTES = I8_BM * 128, I8_B_BYTES = I8_BN * 128;
constexpr size_t I8_SMEM = 2 * I8_S * I8_B_BYTES + I8_ASTAGES * I8_A_BYTES + 1024;
constexpr int I8_THREADS = 320;          // producer warp, MMA warp, 8 epilogue warps

// ---- FP64 panel rows -> 8 int8 slices + scale.  One warp per row, 16 consecutive columns per lane. ----
__global__ void __launch_bounds__(256) slice_panel_kernel(const double* __restrict__ L, int64_t ld, int row0, int nrows, int col0, int ncols,
                                                          int8_t* __restrict__ Sl, double* __restrict__ Se, int64_t cap) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nrows) return;
  const int row = row0 + warp;
  const bool live = 16 * lane < ncols;                        // the panel is 128 * nkb <= 512 columns wide
  const double* src = L + (int64_t)row * ld + col0 + 16 * lane;
  double x[16];
#pragma unroll
  for (int u = 0; u < 16; u += 2) {
    const double2 v = live ? *reinterpret_cast<const double2*>(src + u) : make_double2(0.0, 0.0);
    x[u] = v.x; x[u + 1] = v.y;
  }
  double m = 0.0;
#pragma unroll
  for (int u = 0; u < 16; ++u) m = fmax(m, fabs(x[u]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int e = (m > 0.0) ? ilogb(m) + 1 : 0;                 // |x| < 2^e
  const double sc = ldexp(1.0, 6 - e);                        // |x sc| < 64
  double t[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) t[u] = x[u] * sc;
#pragma unroll
  for (int s = 0; s < I8_S; ++s) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      // digit rule d = floor(t + c), c = 128/255: the remainder stays in [-c, 1 - c], so 256 x remainder + c stays in [-128, 128]
      // and every later digit fits [-128, 127] (the clamp only catches the closed end of that interval); t - d is exact
      const double d = fmin(fmax(floor(t[u] + (128.0 / 255.0)), -128.0), 127.0);
      t[u] = (t[u] - d) * 256.0;
      w[u >> 2] |= ((uint32_t)(uint8_t)(int8_t)(int)d) << (8 * (u & 3));
    }
    if (live) *reinterpret_cast<uint4*>(Sl + ((int64_t)s * cap + row) * I8_K + 16 * lane) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (lane == 0) Se[row] = ldexp(1.0, e - 6);
}

struct I8Maps { CUtensorMap A, B; };   // [slice][row][512 B] with 128-row and 64-row boxes

// tile t of the update: row block bi >= bi_lo, 64-wide column block c2 in [col2_lo, min(col2_hi, 2 bi + 2))
__device__ __forceinline__ void i8_tile(int t, int bi_lo, int col2_lo, int col2_hi, int nblk, int& bi, int& c2) {
  bi = bi_lo; c2 = 0;
  for (; bi < nblk; ++bi) {
    const int hi = (2 * bi + 2 < col2_hi) ? 2 * bi + 2 : col2_hi;
    const int cnt = hi - col2_lo;
    if (cnt > 0) { if (t < cnt) { c2 = col2_lo + t; return; } t -= cnt; }
  }
}

// Persistent: CTA b walks tiles b, b + grid, ...; the pipelines (A ring, B double buffer, TMEM full/empty) run across tiles, so the
// producer prefetches the next tile's operands during the epilogue and the next tile's MMAs start as soon as TMEM has been read.
__global__ void __launch_bounds__(I8_THREADS, 1) syrk_i8_kernel(double* __restrict__ C, int64_t ld, const double* __restrict__ Se, int bi_lo,
                                                                int col2_lo, int col2_hi, int nblk, int ntiles, int nkb,
                                                                const __grid_constant__ I8Maps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;                                     // [2][8 slices][64 rows x 128 B]
  uint8_t* sA = smem_raw + 2 * I8_S * I8_B_BYTES;             // [4][128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + I8_ASTAGES * I8_A_BYTES);
  uint64_t *afull = bars, *aempty = bars + 4, *bfull = bars + 8, *bempty = bars + 10, *tfull = bars + 12, *tempty = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
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
      int as = 0; uint32_t aph = 0, bcnt = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int bi, c2; i8_tile(tile, bi_lo, col2_lo, col2_hi, nblk, bi, c2);
        const int arow = bi * I8_BM, brow = c2 * I8_BN;
        for (int kb = 0; kb < nkb; ++kb, ++bcnt) {
          const int bs = bcnt & 1;
          mbar_wait_or_trap(&bempty[bs], ((bcnt >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bfull[bs], I8_S * I8_B_BYTES);
          for (int q = 0; q < I8_S; ++q) tma_load_3d(sB + (bs * I8_S + q) * I8_B_BYTES, &maps.B, &bfull[bs], kb * 128, brow, q);
          for (int p = 0; p < I8_S; ++p) {
            mbar_wait_or_trap(&aempty[as], aph ^ 1u);
            mbar_arrive_expect_tx(&afull[as], I8_A_BYTES);
            tma_load_3d(sA + as * I8_A_BYTES, &maps.A, &afull[as], kb * 128, arow, p);
            if (++as == I8_ASTAGES) { as = 0; aph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                        // ===== MMA issuer =====
      // D = s32, A = B = signed 8-bit, both K-major, N = 64, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_BN >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
      int as = 0; uint32_t aph = 0, bcnt = 0, it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        mbar_wait_or_trap(tempty, (it & 1u) ^ 1u);            // the epilogue has read the previous tile's accumulators
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int kb = 0; kb < nkb; ++kb, ++bcnt) {
          const int bs = bcnt & 1;
          mbar_wait_or_trap(&bfull[bs], (bcnt >> 1) & 1u);
          for (int p = 0; p < I8_S; ++p) {
            mbar_wait_or_trap(&afull[as], aph);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint64_t da = umma_desc_sw128(smem_u32(sA + as * I8_A_BYTES));
            const uint64_t db = umma_desc_sw128(smem_u32(sB + bs * I8_S * I8_B_BYTES));
            const uint32_t acc = (kb > 0) || (p > 0);         // slice p > 0 always finds its accumulators started by slice p - 1
            // one asm block per slice issues its 4 (7 - p) MMAs with immediate descriptor offsets (umma_issue.cuh)
            switch (p) {
              case 0: umma_i8_issue<0, I8_BN, (I8_B_BYTES >> 4), 1>(tmem, da, db, idesc, acc); break;
              case 1: umma_i8_issue<1, I8_BN, (I8_B_BYTES >> 4), 1>(tmem, da, db, idesc, acc); break;
              case 2: umma_i8_issue<2, I8_BN, (I8_B_BYTES >> 4), 1>(tmem, da, db, idesc, acc); break;
              case 3: umma_i8_issue<3, I8_BN, (I8_B_BYTES >> 4), 1>(tmem, da, db, idesc, acc); break;
              case 4: umma_i8_issue<4, I8_BN, (I8_B_BYTES >> 4), 1>(tmem, da, db, idesc, acc); break;
              case 5: umma_i8_issue<5, I8_BN, (I8_B_BYTES >> 4), 1>(tmem, da, db, idesc, acc); break;
              default: umma_i8_issue<6, I8_BN, (I8_B_BYTES >> 4), 1>(tmem, da, db, idesc, acc); break;
            }
            umma_commit(&aempty[as]);                         // the A stage is free once these MMAs have read it
            if (++as == I8_ASTAGES) { as = 0; aph ^= 1u; }
          }
          umma_commit(&bempty[bs]);
        }
        umma_commit(tfull);                                   // all 448 MMAs of the tile retired: accumulators complete
      }
    }
  } else {
    // ===== epilogue: 8 warps; warp w reads TMEM lanes 32 (w & 3) .. +31 (= rows of the tile), columns 32 half .. +31 =====
    const int g4 = warp & 3, half = (warp - 2) >> 2;
    const int m = 32 * g4 + lane;                             // row inside the tile
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      int bi, c2; i8_tile(tile, bi_lo, col2_lo, col2_hi, nblk, bi, c2);
      const int arow = bi * I8_BM, brow = c2 * I8_BN;
      const double srow = Se[arow + m];
      const int diag_off = arow - brow;                       // element (m, n) is on/below the diagonal iff n <= m + diag_off
      double* crow = C + (int64_t)(arow + m) * ld + brow;
      mbar_wait_or_trap(tfull, it & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // sum_d a_d 2^(-8d) in two exact 64-bit integer groups (|a_d| < 2^26):  (((a0 256 + a1) 256 + a2) 256 + a3) 2^-24 + ((a4 256 + a5) 256 + a6) 2^-48
      double acc[32];
#pragma unroll 1
      for (int grp = 1; grp >= 0; --grp) {                    // the small group first
        long long H[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) H[c] = 0;
        const int nd = grp ? I8_S - 4 : 4;
#pragma unroll 1
        for (int dd = 0; dd < nd; ++dd) {
          uint32_t r[32];
          tmem_ld32(tmem + ((uint32_t)(32 * g4) << 16) + (uint32_t)((4 * grp + dd) * I8_BN + 32 * half), r);
#pragma unroll
          for (int c = 0; c < 32; ++c) H[c] = (H[c] << 8) + (long long)(int32_t)r[c];
        }
        if (grp == 0) {                                       // TMEM has been read: the next tile's MMAs may overwrite it
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
        }
        const double w = grp ? 0x1p-48 : 0x1p-24;
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = grp ? (double)H[c] * w : fma((double)H[c], w, acc[c]);
      }
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const int n = 32 * half + c;
        if (n + 1 <= m + diag_off) {
          const double2 se = *reinterpret_cast<const double2*>(Se + brow + n);
          double2 v = *reinterpret_cast<double2*>(crow + n);
          v.x -= acc[c] * srow * se.x;
          v.y -= acc[c + 1] * srow * se.y;
          *reinterpret_cast<double2*>(crow + n) = v;
        } else if (n <= m + diag_off) {
          crow[n] -= acc[c] * srow * Se[brow + n];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------------------
// Wide variant: 128 x 128 tiles, N = 128 MMAs (64 cycles per 128 x 128 x 32 MMA: twice the work of the 64-wide tile per cycle, whose
// MMAs are bound by re-reading the A operand from shared memory).  Seven 128-column accumulators do not fit the 512 TMEM columns, so
// a tile takes TWO passes over the operands: pass 0 the anti-diagonals d = 0..3 (10 slice products), pass 1 d = 4..6 (18 products).
// Both operands stream: A slices through a 4-deep ring, B slices through one shared-memory slot per slice with its own full/empty
// mbarrier pair -- pass 1 walks p = 6..0, so slot q is released after p = max(0, 4 - q) and refilled for the next k-block while the
// remaining products of this one still run.  The accumulators of a pass are read back (and subtracted from the tile in FP64) while
// the producer already prefetches the next pass.
// MEASURED (N = 8192): 75-79 k cycles per 128 x 128 tile, i.e. no faster than two 64-wide tiles.  At the int8 rate an SS-mode
// 128 x 128 x 32 MMA needs its 8 KB of operands in 64 cycles = the whole 128 B/clk shared-memory port, so the TMA writes into shared
// memory and the MMA reads contend (94 instead of 64 cycles per MMA; MMA-warp waits per tile: 14 k on A stages, 7.5 k on B slots,
// 10-14 k on the TMEM hand-over).  Kept selectable (engine 2) as the starting point for a 2-CTA / A-in-TMEM variant; the 64-wide
// kernel stays the default.
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int I8W_BN = 128;
constexpr uint32_t I8W_T_BYTES = 128 * 128;                         // one slice tile of either operand
constexpr size_t I8W_SMEM = (size_t)(I8_S + I8_ASTAGES) * I8W_T_BYTES + 1024;

__device__ __forceinline__ void i8w_tile(int t, int bi_lo, int cj_lo, int cj_hi, int nblk, int& bi, int& cj) {
  bi = bi_lo; cj = 0;
  for (; bi < nblk; ++bi) {
    const int hi = (bi + 1 < cj_hi) ? bi + 1 : cj_hi;
    const int cnt = hi - cj_lo;
    if (cnt > 0) { if (t < cnt) { cj = cj_lo + t; return; } t -= cnt; }
  }
}
// pass 0: d in [0, 3], p ascending 0..3;  pass 1: d in [4, 6], p descending 6..0
__device__ __forceinline__ int i8w_np(int pass) { return pass ? 7 : 4; }
__device__ __forceinline__ int i8w_p(int pass, int i) { return pass ? 6 - i : i; }
__device__ __forceinline__ int i8w_qlo(int pass, int p) { return pass ? (4 - p > 0 ? 4 - p : 0) : 0; }
__device__ __forceinline__ int i8w_qhi(int pass, int p) { return pass ? 6 - p : 3 - p; }
__device__ __forceinline__ int i8w_last_p(int pass, int q) { return pass ? (4 - q > 0 ? 4 - q : 0) : 3 - q; }   // last p that uses B slice q

__global__ void __launch_bounds__(I8_THREADS, 1) syrk_i8w_kernel(double* __restrict__ C, int64_t ld, const double* __restrict__ Se, int bi_lo,
                                                                 int cj_lo, int cj_hi, int nblk, int ntiles, int nkb,
                                                                 const __grid_constant__ CUtensorMap map) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;                                     // [7 slots][128 rows x 128 B], slot q = slice q of the current k-block
  uint8_t* sA = smem_raw + I8_S * I8W_T_BYTES;                // [4][128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + I8_ASTAGES * I8W_T_BYTES);
  uint64_t *afull = bars, *aempty = bars + 4, *bfull = bars + 8, *bempty = bars + 15, *tfull = bars + 22, *tempty = bars + 23;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < I8_S; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 8);
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
      tma_prefetch_desc(&map);
      int as = 0; uint32_t aph = 0, bgen = 0;                 // bgen: bit q = parity of the number of loads into slot q so far
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int bi, cj; i8w_tile(tile, bi_lo, cj_lo, cj_hi, nblk, bi, cj);
        const int arow = bi * I8_BM, brow = cj * I8W_BN;
        for (int pass = 0; pass < 2; ++pass)
          for (int kb = 0; kb < nkb; ++kb) {
            uint32_t have = 0;                                // B slices already requested for this (pass, k-block)
            for (int i = 0; i < i8w_np(pass); ++i) {
              const int p = i8w_p(pass, i);
              for (int q = i8w_qlo(pass, p); q <= i8w_qhi(pass, p); ++q)
                if (!((have >> q) & 1u)) {
                  mbar_wait_or_trap(&bempty[q], ((bgen >> q) & 1u) ^ 1u);
                  mbar_arrive_expect_tx(&bfull[q], I8W_T_BYTES);
                  tma_load_3d(sB + q * I8W_T_BYTES, &map, &bfull[q], kb * 128, brow, q);
                  have |= 1u << q; bgen ^= 1u << q;
                }
              mbar_wait_or_trap(&aempty[as], aph ^ 1u);
              mbar_arrive_expect_tx(&afull[as], I8W_T_BYTES);
              tma_load_3d(sA + as * I8W_T_BYTES, &map, &afull[as], kb * 128, arow, p);
              if (++as == I8_ASTAGES) { as = 0; aph ^= 1u; }
            }
          }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                        // ===== MMA issuer =====
      // D = s32, A = B = signed 8-bit, both K-major, N = 128, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8W_BN >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
      int as = 0; uint32_t aph = 0, bgen = 0, ev = 0;         // ev: accumulator hand-overs so far (two per tile)
#ifdef I8_PROF
      long long w_t = 0, w_a = 0, w_b = 0, t_all = clock64(), tq; int ntl = 0;
#define I8_TIMED(acc, stmt) do { tq = clock64(); stmt; acc += clock64() - tq; } while (0)
#else
#define I8_TIMED(acc, stmt) do { stmt; } while (0)
#endif
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#ifdef I8_PROF
        ++ntl;
#endif
        for (int pass = 0; pass < 2; ++pass, ++ev) {
          I8_TIMED(w_t, mbar_wait_or_trap(tempty, (ev & 1u) ^ 1u));          // the epilogue has read the previous pass's accumulators
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          uint32_t started = 0;                               // bit a: accumulator a already holds a product of this pass
          const int dmin = pass ? 4 : 0;
          for (int kb = 0; kb < nkb; ++kb) {
            uint32_t have = 0;
            for (int i = 0; i < i8w_np(pass); ++i) {
              const int p = i8w_p(pass, i);
              I8_TIMED(w_a, mbar_wait_or_trap(&afull[as], aph));
              const uint64_t da = umma_desc_sw128(smem_u32(sA + as * I8W_T_BYTES));
              for (int q = i8w_qlo(pass, p); q <= i8w_qhi(pass, p); ++q) {
                if (!((have >> q) & 1u)) {
                  I8_TIMED(w_b, mbar_wait_or_trap(&bfull[q], (bgen >> q) & 1u));
                  have |= 1u << q; bgen ^= 1u << q;
                }
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const int a = p + q - dmin;
                const uint64_t db = umma_desc_sw128(smem_u32(sB + q * I8W_T_BYTES));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_i8(tmem + (uint32_t)(a * I8W_BN), da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc, ((started >> a) & 1u) | (ks > 0));
                started |= 1u << a;
                if (p == i8w_last_p(pass, q)) umma_commit(&bempty[q]);   // slot q may be refilled for the next k-block
              }
              umma_commit(&aempty[as]);
              if (++as == I8_ASTAGES) { as = 0; aph ^= 1u; }
            }
          }
          umma_commit(tfull);                                 // every MMA of the pass retired: accumulators complete
        }
      }
#ifdef I8_PROF
      if (blockIdx.x == 0 && ntl > 0)
        printf("syrk_i8w CTA0: %d tiles, %lld cycles per tile; MMA warp waits per tile: TMEM %lld, A stage %lld, B slot %lld\n", ntl,
               (clock64() - t_all) / ntl, w_t / ntl, w_a / ntl, w_b / ntl);
#endif
    }
  } else {
    // ===== epilogue: 8 warps; warp w reads TMEM lanes 32 (w & 3) .. +31 (= rows of the tile), columns 64 half .. +63 =====
    const int g4 = warp & 3, half = (warp - 2) >> 2;
    const int m = 32 * g4 + lane;
    uint32_t ev = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int bi, cj; i8w_tile(tile, bi_lo, cj_lo, cj_hi, nblk, bi, cj);
      const int arow = bi * I8_BM, brow = cj * I8W_BN;
      const double srow = Se[arow + m];
      const int diag_off = arow - brow;                       // element (m, n) is on/below the diagonal iff n <= m + diag_off
      double* crow = C + (int64_t)(arow + m) * ld + brow;
      for (int pass = 0; pass < 2; ++pass, ++ev) {
        mbar_wait_or_trap(tfull, ev & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int na = pass ? 3 : 4;
        const double w = pass ? 0x1p-48 : 0x1p-24;            // pass 0: ((a0 256 + a1) 256 + a2) 256 + a3;  pass 1: (a4 256 + a5) 256 + a6
        double val[2][32];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          long long H[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) H[c] = 0;
#pragma unroll 1
          for (int a = 0; a < na; ++a) {
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(32 * g4) << 16) + (uint32_t)(a * I8W_BN + 64 * half + 32 * ch), r);
#pragma unroll
            for (int c = 0; c < 32; ++c) H[c] = (H[c] << 8) + (long long)(int32_t)r[c];
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) val[ch][c] = (double)H[c] * w;
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");   // TMEM has been read: the next pass may overwrite it
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const int n = 64 * half + 32 * ch + c;
            if (n + 1 <= m + diag_off) {
              const double2 se = *reinterpret_cast<const double2*>(Se + brow + n);
              double2 v = *reinterpret_cast<double2*>(crow + n);
              v.x -= val[ch][c] * srow * se.x;
              v.y -= val[ch][c + 1] * srow * se.y;
              *reinterpret_cast<double2*>(crow + n) = v;
            } else if (n <= m + diag_off) {
              crow[n] -= val[ch][c] * srow * Se[brow + n];
            }
          }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

cudaError_t make_map3d_u8(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1);

static int tiles_of(int bi_lo, int nblk, int col2_lo, int col2_hi) {
  int n = 0;
  for (int bi = bi_lo; bi < nblk; ++bi) {
    const int hi = (2 * bi + 2 < col2_hi) ? 2 * bi + 2 : col2_hi;
    if (hi > col2_lo) n += hi - col2_lo;
  }
  return n;
}

bool syrk_i8_enabled() {
  static const bool on = !(getenv("B200BO_SYRK_I8") && atoi(getenv("B200BO_SYRK_I8")) == 0);   // B200BO_SYRK_I8=0 selects the DMMA kernel
  return on;
}

// slices of the finished outer panel: rows [row0, Np) x columns [col0, col0 + 512) of the factor
cudaError_t launch_slice_panel(b200bo_handle_s* h, cudaStream_t st, int row0, int col0, int nkb) {
  const int64_t cap = h->cap;
  if (!h->dSl) {
    cudaError_t e = cudaMalloc(&h->dSl, (size_t)I8_S * cap * I8_K);
    if (e == cudaSuccess) e = cudaMalloc(&h->dSe, sizeof(double) * cap);
    if (e == cudaSuccess) e = make_map3d_u8(&h->tmSlA, h->dSl, I8_K, (uint64_t)cap, I8_S, 128, I8_BM);
    if (e == cudaSuccess) e = make_map3d_u8(&h->tmSlB, h->dSl, I8_K, (uint64_t)cap, I8_S, 128, I8_BN);
    if (e != cudaSuccess) return e;
  }
  const int nrows = (int)h->Np - row0;
  if (nrows <= 0) return cudaSuccess;
  slice_panel_kernel<<<(nrows + 7) / 8, 256, 0, st>>>(h->dL, h->ld, row0, nrows, col0, 128 * nkb, reinterpret_cast<int8_t*>(h->dSl), h->dSe, cap);
  h->i8_nkb = nkb;
  h->launches++;
  return cudaGetLastError();
}

// A_ij -= L_i,P L_j,P^T over the tiles (bi >= bi_lo, 64-wide column blocks [col2_lo, min(col2_hi, 2 bi + 2))) from the current slices
cudaError_t launch_syrk_i8(b200bo_handle_s* h, cudaStream_t st, int bi_lo, int col2_lo, int col2_hi, int* ntiles, int max_ctas) {
  const int nblk = (int)(h->Np / NB);
  static const bool wide_env = getenv("B200BO_SYRK_I8") && atoi(getenv("B200BO_SYRK_I8")) == 2;
  if (h->syrk_engine == 2 || (h->syrk_engine < 0 && wide_env)) {   // 128 x 128 tiles, two passes (measured: no faster, see the header)
    const int cj_lo = col2_lo / 2, cj_hi = col2_hi / 2;         // both limits are multiples of a 128-column block
    int n = 0;
    for (int bi = bi_lo; bi < nblk; ++bi) { const int hi = (bi + 1 < cj_hi) ? bi + 1 : cj_hi; if (hi > cj_lo) n += hi - cj_lo; }
    if (ntiles) *ntiles = n;
    if (n == 0) return cudaSuccess;
    cudaFuncSetAttribute(syrk_i8w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8W_SMEM);
    const int grid = n < max_ctas ? n : max_ctas;
    syrk_i8w_kernel<<<grid, I8_THREADS, I8W_SMEM, st>>>(h->dL, h->ld, h->dSe, bi_lo, cj_lo, cj_hi, nblk, n, h->i8_nkb, h->tmSlA);
    h->launches++;
    return cudaGetLastError();
  }
  const int n = tiles_of(bi_lo, nblk, col2_lo, col2_hi);
  if (ntiles) *ntiles = n;
  if (n == 0) return cudaSuccess;
  cudaFuncSetAttribute(syrk_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8_SMEM);
  I8Maps maps;
  maps.A = h->tmSlA; maps.B = h->tmSlB;
  const int grid = n < max_ctas ? n : max_ctas;
  syrk_i8_kernel<<<grid, I8_THREADS, I8_SMEM, st>>>(h->dL, h->ld, h->dSe, bi_lo, col2_lo, col2_hi, nblk, n, h->i8_nkb, maps);
  h->launches++;
  return cudaGetLastError();
}

}  // namespace b200bo
