// tma.cuh -- Blackwell async-copy primitives: TMA tile loads (cp.async.bulk.tensor, SASS UTMALDG) into 128B-swizzled
// shared-memory tiles, mbarrier full/empty pipelines, and the DMMA fragment loads that match the TMA swizzle.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace b200bo {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// writes made through the generic proxy (st.global) become visible to later async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }

// ---- TMA 2D tile load: box (c0 = inner/k element index, c1 = row index) -> smem, completes on `bar` -------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- register redistribution between the producer and consumer warpgroups ---------------------------------------
template <int R> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(R)); }

// named barrier over the consumer warps only (the producer warp never joins)
template <int NTHREADS> __device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(NTHREADS) : "memory"); }

// ---- tile addressing ----------------------------------------------------------------------------------------------
// A tile holds rows of 16 doubles (128 B) or, for the resident operand, `stride` doubles; the 16-byte chunk index is
// XORed with (row & 7): exactly CU_TENSOR_MAP_SWIZZLE_128B for 128-byte rows on a 1024-byte aligned tile.
__device__ __forceinline__ int toff(int row, int k, int stride) { return row * stride + ((((k >> 1) ^ (row & 7))) << 1) + (k & 1); }

// MMA row slot g (0..7) is mapped to tile row rho(g) so that the two rows read by one quarter-warp differ in bit 2:
// with the XOR above their four 16-byte chunks land in disjoint bank groups -> LDS.128 fragment loads are conflict-free.
// Consequence for the accumulator: acc[mt][nt][e] of lane (g, q) is element
//     row = warp_row0 + 8 mt + rho(g),   col = warp_col0 + 8 nt + (e ? q + 4 : q)          (rho(2q) = q, rho(2q+1) = q + 4)
__device__ __forceinline__ int rho(int g) { return (g >> 1) | ((g & 1) << 2); }

__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// Per-thread fragment addressing, hoisted out of the chunk loop.  For a tile whose rows are `row_bytes` apart, lane (g, q) of
// a warp whose first row is row0 reads, for k-half kk (0/1) of a 16-wide chunk at chunk-aligned k offset kb (bytes):
//     base + (row0 + 8 t + rho(g)) * row_bytes + kb + xk[kk],      xk[kk] = ((4 kk + q) ^ (rho(g) & 7)) * 16
struct FragAddr {
  uint32_t xk[2];
  __device__ __forceinline__ FragAddr(int rg, int q) { xk[0] = (uint32_t)((q ^ (rg & 7)) << 4); xk[1] = (uint32_t)(((4 + q) ^ (rg & 7)) << 4); }
};

// acc += A(MT*8 rows x 16 k) * B(NT*8 rows x 16 k)^T for one warp.  aaddr/baddr: shared-space byte address of this lane's row
// in the first row tile, INCLUDING the chunk's k offset; consecutive row tiles are 8 * row_bytes apart.
template <int MT, int NT, int A_ROW_BYTES, int B_ROW_BYTES>
__device__ __forceinline__ void warp_mma_chunk_t(double (&acc)[MT][NT][2], uint32_t aaddr, uint32_t baddr, const FragAddr& f) {
#pragma unroll
  for (int kk = 0; kk < KC / 8; ++kk) {
    double2 a[MT], b[NT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) a[mt] = lds128(aaddr + f.xk[kk] + mt * 8 * A_ROW_BYTES);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b[nt] = lds128(baddr + f.xk[kk] + nt * 8 * B_ROW_BYTES);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
  }
}

struct PipeState {
  int s = 0;
  uint32_t ph = 0;
  template <int NS> __device__ __forceinline__ void next() {
    if (++s == NS) { s = 0; ph ^= 1u; }
  }
};

}  // namespace b200bo
