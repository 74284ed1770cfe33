// umma.cuh -- tcgen05 plumbing shared by the int8-slice kernels (syrk_i8.cu: K4, acq_i8.cu: K6): shared-memory matrix descriptors for
// SWIZZLE_128B K-major tiles, tcgen05.mma kind::i8 issue, tcgen05.commit -> mbarrier, bounded mbarrier waits (trap instead of hanging
// the GPU), 3-D TMA tile loads and the 32-lane x 32-column TMEM read-back.
#pragma once
#include "tma.cuh"

namespace b200bo {

// One lane of a CONVERGED warp (elect.sync).  Unlike `lane == 0`, ptxas can prove that a region guarded by this predicate runs on a
// single thread, so every tcgen05.mma / cp.async.bulk.tensor inside compiles to one UTCIMMA / UTMALDG; behind `lane == 0` each of
// them is wrapped in an election loop (ELECT / PLOP3 / BRA.U.ANY, ~8 dependent scalar instructions = 45-60 cycles per 32-cycle MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
  return pred != 0;
}

// ---- tcgen05 plumbing ----
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {   // K-major, SWIZZLE_128B, 8-row groups 1024 B apart (version 1)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int i = 0; i < (1 << 24) && !ok; ++i)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (!ok) __trap();
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}


// Balanced radix-256 digits of a 55-bit signed integer q (|q| <= 2^54):  q = d0 2^48 + sum_{p=1..6} d_p 2^(8 (6 - p)),  d_p in [-128, 127]
// (p >= 1), d0 in [-64, 64].  Adding the bias  sum_{p>=1} 128 * 2^(8 (6 - p))  makes every lower digit an unsigned byte u_p = d_p + 128 of
// the biased integer; flipping the top bit of that byte (u ^ 0x80) is d_p as a two's-complement int8.  Byte k of the result = slice 6 - k.
__device__ __forceinline__ unsigned long long i8_digits(long long q) {
  const unsigned long long biased = (unsigned long long)(q + 0x0000808080808080LL);
  return biased ^ 0x0000808080808080ULL;
}

}  // namespace b200bo
