// tilegemm.cuh -- the TMA + DMMA tile GEMM shared by the factorisation (chol.cu: K3, K4) and the symmetric inverse (kinv.cu).
#pragma once
#include "tma.cuh"

namespace b200bo {

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

}  // namespace b200bo
