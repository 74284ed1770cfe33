// peak.cu -- self-measured FP64 tensor-pipe denominator: back-to-back independent DMMA.8x8x4 chains on every SM.
// MEASURED_PEAKS.json carries only HBM and bf16 figures; the path computes in FP64, so the roofline of the
// contraction kernels (K4, K6) is quoted against this number (measured 37.1 TFLOP/s on the pool's B200s).
#include "common.cuh"
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

}  // namespace b200bo
