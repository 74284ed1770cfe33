// Self-measured FP64 denominators on B200: DFMA (FMA pipe) and DMMA.8x8x4 (tensor pipe) peak rates.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters) {
  double a[8], b = 1.0000001, c = 0.5;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2]; double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {256, 512, 1024}) {
    for (int rep = 0; rep < 2; ++rep) {
      const int iters = 20000, blocks = p.multiProcessorCount * (1024 / threads);
      float ms;
      cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      double fl = 2.0 * 8 * iters * (double)blocks * threads;
      if (rep) printf("DFMA  threads=%4d blocks=%4d  %.3f ms  %.2f TFLOP/s\n", threads, blocks, ms, fl / ms * 1e-9);
      cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      fl = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32);
      if (rep) printf("DMMA  threads=%4d blocks=%4d  %.3f ms  %.2f TFLOP/s\n", threads, blocks, ms, fl / ms * 1e-9);
    }
  }
  printf("SMs=%d clock=%d kHz\n", p.multiProcessorCount, p.clockRate);
  return 0;
}
