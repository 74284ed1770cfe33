// Developer probe: store-only HBM bandwidth on this GPU (the practical ceiling for K1, which only writes).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill256(double* p, size_t n4, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.v4.f64 [%0], {%1,%1,%1,%1};" ::"l"(p + 4 * i), "d"(v) : "memory");
}
__global__ void fill128(double2* p, size_t n2, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) p[i] = make_double2(v, v);
}
__global__ void copy128(const double2* __restrict__ a, double2* __restrict__ b, size_t n2) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
int main() {
  for (size_t N : {4096ul, 8192ul}) {
    size_t bytes = N * N * 8;
    double *d, *e, *flush; cudaMalloc(&d, bytes); cudaMalloc(&e, bytes); cudaMalloc(&flush, 256u << 20);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 4; ++mode) {
      float best = 1e9, sum = 0;
      for (int it = 0; it < 8; ++it) {
        cudaMemsetAsync(flush, 0, 256u << 20);
        cudaEventRecord(a);
        if (mode == 0) fill256<<<148 * 8, 256>>>(d, bytes / 32, 1.0);
        else if (mode == 1) fill128<<<148 * 8, 256>>>((double2*)d, bytes / 16, 1.0);
        else if (mode == 2) cudaMemsetAsync(d, 0, bytes);
        else copy128<<<148 * 8, 256>>>((const double2*)d, (double2*)e, bytes / 16);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
      }
      const char* nm[] = {"fill256", "fill128", "memset", "copy(r+w)"};
      double mv = (mode == 3 ? 2.0 : 1.0) * bytes;
      printf("N=%zu %s: best %.1f us (%.0f GB/s) avg %.1f us (%.0f GB/s)\n", N, nm[mode], best * 1e3, mv / best * 1e-6, sum / 6 * 1e3, mv / (sum / 6) * 1e-6);
    }
    cudaFree(d); cudaFree(e); cudaFree(flush);
  }
  return 0;
}
