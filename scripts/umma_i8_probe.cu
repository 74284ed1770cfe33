// Developer probe (round-2 groundwork): tcgen05.mma kind::i8 on sm_100a -- descriptor layout check against a CPU product and the
// int8 tensor rate of this GPU, i.e. the denominator an Ozaki-slice FP64 emulation of the trailing update would be quoted against.
//   C[128 x N] (s32, TMEM) = A[128 x K] (s8, K-major, 128B swizzle) * B[N x K]^T (s8, K-major, 128B swizzle)
// One CTA = 128 threads.  Operands are written into shared memory in the canonical 8-row x 128-byte swizzle atoms by the threads
// themselves (so the check does not depend on a tensor map), one elected thread issues the MMAs, completion arrives on an mbarrier
// through tcgen05.commit, four warps read the accumulator back with tcgen05.ld.  Every wait is bounded: a failure sets a flag
// instead of hanging the GPU.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {     // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address
  d |= (uint64_t)0 << 16;                            // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int spins) {
  uint32_t ok = 0;
  for (int i = 0; i < spins && !ok; ++i)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <int N>
__global__ void __launch_bounds__(128, 1) umma_i8_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ C, int K,
                                                         int reps, int* __restrict__ err, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = K / 128;                           // k-blocks of 128 bytes
  uint8_t* sA = smem;                                // [nkb][128 rows][128 B]
  uint8_t* sB = smem + (size_t)nkb * 128 * 128;      // [nkb][N rows][128 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // operands -> canonical swizzled tiles: 16-byte chunk c of row r lands at (r/8)*1024 + (r%8)*128 + ((c ^ (r%8)) * 16)
  for (int e = tid; e < nkb * 128 * 8; e += 128) {
    const int kb = e / (128 * 8), r = (e / 8) % 128, c = e % 8;
    const int4 v = *reinterpret_cast<const int4*>(A + (size_t)r * K + kb * 128 + c * 16);
    *reinterpret_cast<int4*>(sA + (size_t)kb * 128 * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  for (int e = tid; e < nkb * N * 8; e += 128) {
    const int kb = e / (N * 8), r = (e / 8) % N, c = e % 8;
    const int4 v = *reinterpret_cast<const int4*>(B + (size_t)r * K + kb * 128 + c * 16);
    *reinterpret_cast<int4*>(sB + (size_t)kb * N * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {                                   // TMEM: N columns of 32-bit accumulators (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy operand writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      for (int kb = 0; kb < nkb; ++kb) {
        const uint64_t da = make_desc(smem_u32(sA + (size_t)kb * 128 * 128)), db = make_desc(smem_u32(sB + (size_t)kb * N * 128));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {             // UMMA_K = 32 bytes: advance the start address by 32 B inside the swizzle atom
          const uint32_t acc = (rep | kb | ks) != 0;
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem),
              "l"(da + (uint64_t)(ks * 2)), "l"(db + (uint64_t)(ks * 2)), "r"(idesc), "r"(acc), "r"(0u)
              : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
  }
  const bool done = mbar_wait_bounded(&bar, 0, 1 << 22);
  if (tid == 0 && cycles) cycles[blockIdx.x] = clock64() - t0;
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (!done) { if (tid == 0) atomicExch(err, 1); }
  if (done && C && blockIdx.x == 0) {
    // accumulator read-back: warp w owns TMEM lanes 32 w .. 32 w + 31 (= rows of D), 32 columns per tcgen05.ld
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
            "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
            "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
            "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) C[(size_t)(32 * warp + lane) * N + c0 + i] = (int32_t)r[i];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((uint32_t)N) : "memory");
}

template <int N>
static int run(int K) {
  std::vector<int8_t> hA((size_t)128 * K), hB((size_t)N * K);
  srand(1234 + N);
  for (auto& v : hA) v = (int8_t)(rand() % 129 - 64);
  for (auto& v : hB) v = (int8_t)(rand() % 129 - 64);
  int8_t *dA, *dB; int32_t* dC; int* derr; long long* dcyc;
  cudaMalloc(&dA, hA.size()); cudaMalloc(&dB, hB.size()); cudaMalloc(&dC, (size_t)128 * N * 4); cudaMalloc(&derr, 4); cudaMalloc(&dcyc, 148 * 8);
  cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
  cudaMemset(derr, 0, 4); cudaMemset(dC, 0xff, (size_t)128 * N * 4);
  const size_t smem = (size_t)(K / 128) * (128 + N) * 128;
  cudaFuncSetAttribute(umma_i8_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_i8_kernel<N><<<1, 128, smem>>>(dA, dB, dC, K, 1, derr, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0; cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  std::vector<int32_t> hC((size_t)128 * N);
  cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      int32_t s = 0;
      for (int k = 0; k < K; ++k) s += (int32_t)hA[(size_t)m * K + k] * (int32_t)hB[(size_t)n * K + k];
      if (s != hC[(size_t)m * N + n]) { if (bad < 4) printf("  mismatch (%d,%d): got %d want %d\n", m, n, hC[(size_t)m * N + n], s); ++bad; }
    }
  printf("N=%d K=%d: cuda=%s timeout=%d mismatches=%ld of %d\n", N, K, cudaGetErrorString(e), herr, bad, 128 * N);
  if (e == cudaSuccess && !herr && bad == 0) {       // rate: every SM issues reps x K/32 MMAs of 128 x N x 32 on resident operands
    const int reps = 2000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    umma_i8_kernel<N><<<148, 128, smem>>>(dA, dB, nullptr, K, reps, derr, dcyc);
    cudaEventRecord(a);
    umma_i8_kernel<N><<<148, 128, smem>>>(dA, dB, nullptr, K, reps, derr, dcyc);
    cudaEventRecord(b);
    e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    long long cyc[148]; cudaMemcpy(cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    const double macs = 148.0 * reps * (double)K * 128.0 * N;
    printf("  rate: %.3f ms, %.1f int8 TOPS (2 ops per MAC) over 148 SMs, %.0f MAC/clk/SM by clock64 (cuda=%s timeout=%d)\n", ms, 2.0 * macs / (ms * 1e-3) * 1e-12,
           (double)reps * K * 128.0 * N / (double)cyc[0], cudaGetErrorString(e), herr);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(derr); cudaFree(dcyc);
  return (e == cudaSuccess && !herr && bad == 0) ? 0 : 1;
}

int main() {
  int rc = 0;
  rc |= run<64>(256);
  rc |= run<128>(256);
  rc |= run<256>(256);
  return rc;
}
