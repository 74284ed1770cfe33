// Developer probe: tcgen05.mma kind::i8 with the A operand in TENSOR MEMORY (TS mode), staged there by tcgen05.cp from the same
// SWIZZLE_128B K-major shared-memory tile the SS-mode MMAs read.  Why: the 128 x 64 x 32 slice MMAs of K4 / K6 re-read their 4 KB A
// chunk from shared memory for every one of up to 7 B slices, and the shared-memory port (128 B/clk), not the tensor pipe, bounds them
// (840 KB of port traffic per k-block of the 7 x 7 slice product).  With A in TMEM an MMA reads only its 2 KB B chunk.
//   1. tcgen05.cp.128x256b of the four 32-byte k-steps of a 128-row tile -> 32 TMEM columns; read back and compared byte for byte
//   2. D[128 x 64] (s32) = A[tmem] * B[smem]^T compared with the CPU product
//   3. issue rates: SS, TS, and the real mix (4 copies + 4 (7 - p) MMAs per slice p)
// Every wait is bounded (a failure sets a flag instead of hanging the GPU).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {     // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int spins) {
  uint32_t ok = 0;
  for (int i = 0; i < spins && !ok; ++i)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// SS MMA with an A-collector hint: 0 = fill (read A from shared memory and keep it), 1 = use (reuse the kept A, keep it), 2 = lastuse
template <int C>
__device__ __forceinline__ void mma_ss_c(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (C == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else if (C == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
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

constexpr int N = 64;
// mode 0: correctness (copy, read back A, TS product, read back D);  1: SS rate;  2: TS rate;  3: real mix per slice (4 cp + 4 (7-p) TS MMAs)
__global__ void __launch_bounds__(128, 1) probe(const int8_t* __restrict__ A, const int8_t* __restrict__ B, uint32_t* __restrict__ Aback,
                                                int32_t* __restrict__ C, int mode, int reps, int* __restrict__ err, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                     // [128 rows][128 B]
  uint8_t* sB = smem + 128 * 128;         // [7][64 rows][128 B] (the same tile seven times in the rate modes)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 128 * 8; e += 128) {
    const int r = e / 8, c = e % 8;
    *reinterpret_cast<int4*>(sA + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const int4*>(A + (size_t)r * 128 + c * 16);
  }
  for (int e = tid; e < 7 * N * 8; e += 128) {
    const int q = e / (N * 8), r = (e / 8) % N, c = e % 8;
    int4 v = *reinterpret_cast<const int4*>(B + (size_t)r * 128 + c * 16);
    int8_t* vb = reinterpret_cast<int8_t*>(&v);
    for (int t = 0; t < 16; ++t) vb[t] = (int8_t)(vb[t] + q);       // slice q = B + q: seven different matrices
    *reinterpret_cast<int4*>(sB + q * N * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tA = tmem + 448;                    // two 32-column A buffers behind the seven 64-column accumulators
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
  long long t0 = 0;
  uint32_t elected = 0;
  if (warp == 0) asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(elected));
  if (elected) {            // elect.sync, not `tid == 0`: ptxas then emits plain back-to-back UTCIMMA (no per-MMA election loop)
    t0 = clock64();
    if (mode == 0) {
      for (int ks = 0; ks < 4; ++ks) cp_128x256b(tA + 8 * ks, da + 2 * ks);
      for (int ks = 0; ks < 4; ++ks) mma_ts(tmem, tA + 8 * ks, db + 2 * ks, idesc, ks > 0);
    } else if (mode == 1) {
      for (int rep = 0; rep < reps; ++rep)
        for (int q = 0; q < 7; ++q)
          for (int ks = 0; ks < 4; ++ks) mma_ss(tmem + 64 * q, da + 2 * ks, db + q * 512 + 2 * ks, idesc, 1);
    } else if (mode == 2) {
      for (int ks = 0; ks < 4; ++ks) cp_128x256b(tA + 8 * ks, da + 2 * ks);
      for (int rep = 0; rep < reps; ++rep)
        for (int q = 0; q < 7; ++q)
          for (int ks = 0; ks < 4; ++ks) mma_ts(tmem + 64 * q, tA + 8 * ks, db + q * 512 + 2 * ks, idesc, 1);
    } else if (mode == 4 || mode == 5) {
      // k-step outer, B slice inner: the A chunk of a k-step is read once (fill), reused by the next five MMAs, released by the seventh
      for (int rep = 0; rep < reps; ++rep)
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acc = mode == 4 ? 1u : (uint32_t)(ks > 0);
          mma_ss_c<0>(tmem, da + 2 * ks, db + 2 * ks, idesc, acc);
          for (int q = 1; q < 6; ++q) mma_ss_c<1>(tmem + 64 * q, da + 2 * ks, db + q * 512 + 2 * ks, idesc, acc);
          mma_ss_c<2>(tmem + 64 * 6, da + 2 * ks, db + 6 * 512 + 2 * ks, idesc, acc);
        }
    } else if (mode == 6) {
      for (int rep = 0; rep < reps; ++rep)
        asm volatile("{\n"
        ".reg .pred pt;\n.reg .b64 ta, tb;\n.reg .b32 td;\nsetp.eq.b32 pt, 0, 0;\n"
        "add.s32 td, %0, 0;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 0;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::fill [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 64;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 512;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 128;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 1024;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 192;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 1536;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 256;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 2048;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 320;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 2560;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 384;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 3072;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 0;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 2;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::fill [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 64;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 514;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 128;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 1026;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 192;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 1538;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 256;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 2050;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 320;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 2562;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 384;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 3074;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 0;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 4;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::fill [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 64;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 516;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 128;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 1028;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 192;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 1540;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 256;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 2052;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 320;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 2564;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 384;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 3076;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 0;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 6;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::fill [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 64;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 518;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 128;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 1030;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 192;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 1542;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 256;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 2054;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 320;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 2566;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 384;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 3078;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [td], ta, tb, %3, pt;\n"
        "}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc) : "memory");
    } else if (mode == 7) {
      for (int rep = 0; rep < reps; ++rep)
        asm volatile("{\n"
        ".reg .pred pt;\n.reg .b64 ta, tb;\n.reg .b32 td;\nsetp.eq.b32 pt, 0, 0;\n"
        "add.s32 td, %0, 0;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 0;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 64;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 512;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 128;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 1024;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 192;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 1536;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 256;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 2048;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 320;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 2560;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 384;\nadd.s64 ta, %1, 0;\nadd.s64 tb, %2, 3072;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 0;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 2;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 64;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 514;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 128;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 1026;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 192;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 1538;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 256;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 2050;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 320;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 2562;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 384;\nadd.s64 ta, %1, 2;\nadd.s64 tb, %2, 3074;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 0;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 4;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 64;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 516;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 128;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 1028;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 192;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 1540;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 256;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 2052;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 320;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 2564;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 384;\nadd.s64 ta, %1, 4;\nadd.s64 tb, %2, 3076;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 0;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 6;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 64;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 518;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 128;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 1030;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 192;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 1542;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 256;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 2054;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 320;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 2566;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "add.s32 td, %0, 384;\nadd.s64 ta, %1, 6;\nadd.s64 tb, %2, 3078;\ntcgen05.mma.cta_group::1.kind::i8 [td], ta, tb, %3, pt;\n"
        "}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc) : "memory");
    } else {
      for (int rep = 0; rep < reps; ++rep)
        for (int p = 0; p < 7; ++p) {
          const uint32_t ta = tA + 32 * (p & 1);
          for (int ks = 0; ks < 4; ++ks) cp_128x256b(ta + 8 * ks, da + 2 * ks);
          for (int q = 0; q + p < 7; ++q)
            for (int ks = 0; ks < 4; ++ks) mma_ts(tmem + 64 * (p + q), ta + 8 * ks, db + q * 512 + 2 * ks, idesc, 1);
        }
    }
    commit(&bar);
  }
  const bool done = mbar_wait_bounded(&bar, 0, 1 << 24);
  if (elected && cycles) cycles[blockIdx.x] = clock64() - t0;
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (!done && tid == 0) atomicExch(err, 1);
  if (done && mode == 5 && blockIdx.x == 0) {
    uint32_t r[32];
    for (int c0 = 0; c0 < 448; c0 += 32) {
      tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
      for (int i = 0; i < 32; ++i) C[(size_t)(32 * warp + lane) * 448 + c0 + i] = (int32_t)r[i];
    }
  }
  if (done && mode == 0 && blockIdx.x == 0) {
    uint32_t r[32];
    tmem_ld32(tA + ((uint32_t)(32 * warp) << 16), r);
    for (int i = 0; i < 32; ++i) Aback[(size_t)(32 * warp + lane) * 32 + i] = r[i];
    for (int c0 = 0; c0 < N; c0 += 32) {
      tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
      for (int i = 0; i < 32; ++i) C[(size_t)(32 * warp + lane) * N + c0 + i] = (int32_t)r[i];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  std::vector<int8_t> hA(128 * 128), hB(N * 128);
  srand(99);
  for (auto& v : hA) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : hB) v = (int8_t)(rand() % 256 - 128);
  int8_t *dA, *dB; uint32_t* dAb; int32_t* dC; int* derr; long long* dcyc;
  cudaMalloc(&dA, hA.size()); cudaMalloc(&dB, hB.size()); cudaMalloc(&dAb, 128 * 32 * 4); cudaMalloc(&dC, 128 * 448 * 4); cudaMalloc(&derr, 4); cudaMalloc(&dcyc, 148 * 8);
  cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
  cudaMemset(derr, 0, 4); cudaMemset(dC, 0xff, 128 * N * 4); cudaMemset(dAb, 0xee, 128 * 32 * 4);
  const size_t smem = 128 * 128 + 7 * N * 128;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dA, dB, dAb, dC, 0, 1, derr, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0; cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  std::vector<uint32_t> hAb(128 * 32); std::vector<int32_t> hC(128 * N);
  cudaMemcpy(hAb.data(), dAb, hAb.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
  long badA = 0, badC = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 32; ++c) {
      uint32_t want = 0;
      for (int b = 0; b < 4; ++b) want |= (uint32_t)(uint8_t)hA[r * 128 + 4 * c + b] << (8 * b);
      if (want != hAb[r * 32 + c]) { if (badA < 6) printf("  A readback lane %d col %d: got %08x want %08x\n", r, c, hAb[r * 32 + c], want); ++badA; }
    }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      int32_t s = 0;
      for (int k = 0; k < 128; ++k) s += (int32_t)hA[m * 128 + k] * (int32_t)hB[n * 128 + k];
      if (s != hC[m * N + n]) { if (badC < 4) printf("  D mismatch (%d,%d): got %d want %d\n", m, n, hC[m * N + n], s); ++badC; }
    }
  printf("TS correctness: cuda=%s timeout=%d  A-in-TMEM mismatches=%ld of 4096  D mismatches=%ld of %d\n", cudaGetErrorString(e), herr, badA, badC, 128 * N);
  if (e != cudaSuccess || herr) return 1;
  {
    // collector correctness: seven accumulators, each must hold A B_q^T with B_q = B + q (a different matrix per slice)
    probe<<<1, 128, smem>>>(dA, dB, dAb, dC, 5, 1, derr, nullptr);
    e = cudaDeviceSynchronize();
    std::vector<int32_t> hC7(128 * 448);
    cudaMemcpy(hC7.data(), dC, hC7.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int q = 0; q < 7; ++q)
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          int32_t sacc = 0;
          for (int k = 0; k < 128; ++k) sacc += (int32_t)hA[m * 128 + k] * (int32_t)(int8_t)(hB[n * 128 + k] + q);
          if (sacc != hC7[m * 448 + 64 * q + n]) { if (bad < 4) printf("  collector mismatch q=%d (%d,%d): got %d want %d\n", q, m, n, hC7[m * 448 + 64 * q + n], sacc); ++bad; }
        }
    printf("A-collector (fill/use/lastuse) correctness: cuda=%s timeout=%d mismatches=%ld of %d\n", cudaGetErrorString(e), herr, bad, 7 * 128 * N);
  }
  const char* names[8] = {"", "SS  (A, B from smem)", "TS  (A from TMEM)", "mix (per slice: 4 cp + 4(7-p) TS MMAs)", "SS + A collector (fill, 5 x use, lastuse)", "",
                          "SS + A collector, one asm block of 28 MMAs", "SS, one asm block of 28 MMAs"};
  for (int mode = 1; mode <= 7; ++mode) {
    if (mode == 5) continue;
    const int reps = 400;
    probe<<<148, 128, smem>>>(dA, dB, dAb, dC, mode, reps, derr, dcyc);
    probe<<<148, 128, smem>>>(dA, dB, dAb, dC, mode, reps, derr, dcyc);
    e = cudaDeviceSynchronize();
    long long cyc[148]; cudaMemcpy(cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    const double mmas = (double)reps * (mode == 3 ? 112.0 : 28.0);   // modes 1, 2, 4: 7 x 4 per rep
    printf("%-40s: %.1f cycles per 128x64x32 MMA (%.0f MMAs, cuda=%s timeout=%d)\n", names[mode], (double)cyc[0] / mmas, mmas, cudaGetErrorString(e), herr);
  }
  return 0;
}
