// peak.cu -- self-measured FP64 tensor-pipe denominator: back-to-back independent DMMA.8x8x4 chains on every SM.
// MEASURED_PEAKS.json carries only HBM and bf16 figures; the path computes in FP64, so the roofline of the
// contraction kernels (K4, K6) is quoted against this number (measured 37.1 TFLOP/s on the pool's B200s).
#include "umma.cuh"
#include "handle.h"

namespace b200bo {

__global__ void __launch_bounds__(512) dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
  const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t launch_dmma_peak(b200bo_handle_s* h, double* tflops) {
  const int iters = 20000, threads = 512, blocks = h->num_sms * 2;
  double* out = h->dV;                 // scratch, >= blocks*threads doubles
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(h->ev[6], h->stream);
    dmma_peak_kernel<<<blocks, threads, 0, h->stream>>>(out, iters);
    cudaEventRecord(h->ev[7], h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return e;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]);
    if (rep > 0 && ms < best) best = ms;
  }
  *tflops = 2.0 * 256.0 * 8.0 * iters * (double)blocks * (threads / 32) / (best * 1e-3) * 1e-12;
  return cudaGetLastError();
}

// ---- self-measured int8 tensor-pipe denominator (tcgen05.mma kind::i8): the roofline of the int8-slice kernels (K4 on syrk_i8.cu, K6 on
// acq_i8.cu).  Every SM issues back-to-back 128 x 256 x 32 MMAs on operands resident in shared memory (A 16 KB, B 32 KB, SWIZZLE_128B
// K-major; the values do not matter), alternating between two 256-column accumulators.  MEASURED_PEAKS.json has no int8 figure; the
// nominal dense rate is 4.5 POP/s (twice bf16).
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int reps, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < (16384 + 32768) / 16; e += 128) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0x01010101u * (e & 3), 0x02020202u, 0x01010101u, 0u);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = umma_desc_sw128(smem_u32(smem)), db = umma_desc_sw128(smem_u32(smem + 16384));
    for (int r = 0; r < reps; ++r)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_i8(tmem + 256u * (uint32_t)(ks & 1), da + 2 * ks, db + 2 * ks, idesc, 1u);
    umma_commit(&bar);
  }
  uint32_t ok = 0;
  for (int i = 0; i < (1 << 26) && !ok; ++i)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  if (!ok && tid == 0) atomicExch(err, 1);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

cudaError_t launch_i8_peak(b200bo_handle_s* h, double* tops) {
  const int reps = 20000, blocks = h->num_sms;
  const size_t smem = 16384 + 32768;
  cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaError_t e = cudaMemsetAsync(h->dinfo + 2, 0, sizeof(int), h->stream);
  if (e != cudaSuccess) return e;
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(h->ev[6], h->stream);
    i8_peak_kernel<<<blocks, 128, smem, h->stream>>>(reps, h->dinfo + 2);
    cudaEventRecord(h->ev[7], h->stream);
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return e;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]);
    if (rep > 0 && ms < best) best = ms;
  }
  int bad = 0;
  if ((e = cudaMemcpy(&bad, h->dinfo + 2, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
  if (bad) return cudaErrorLaunchTimeout;
  *tops = 2.0 * 128.0 * 256.0 * 32.0 * 4.0 * reps * (double)blocks / (best * 1e-3) * 1e-12;
  return cudaGetLastError();
}

}  // namespace b200bo
