// Developer probe: single-warp latencies/throughputs that bound the diagonal-block kernel (K2).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double d) {
  double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0); const double t = fma(e, e, e); return fma(y, t, y);
}
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[1024];
  const int lane = threadIdx.x;
  for (int i = lane; i < 1024; i += 32) sm[i] = seed * 1e-3 * i;
  __syncwarp();
  double x = seed, acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = seed + i;
  long long t0, t1;
  // a) dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = fma(x, 1.0000001, 1e-9);
  t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
  // b) 8 independent chains
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) acc[i & 7] = fma(acc[i & 7], 1.0000001, 1e-9);
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
  // c) independent double shuffles + dependent-free FMA
  double s = 0.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) s += __shfl_sync(0xffffffffu, acc[i & 7], i & 31);
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
  // d) rcp chain
  double r = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) r = fast_rcp(r) + 0.5;
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
  // e) broadcast LDS.64 + FMA into 8 accumulators
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) acc[i & 7] = fma(sm[i], x, acc[i & 7]);
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
  // f) shuffle -> fma dependent chain (double)
  double c = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) c = fma(__shfl_sync(0xffffffffu, c, (i + 1) & 31), 0.999, 1e-3);
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
  // g) 32 accumulators (like the kernel): broadcast LDS.128 + 2 FMA
  double a32[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a32[i] = seed * i;
  t0 = clock64();
#pragma unroll
  for (int rep = 0; rep < 8; ++rep)
#pragma unroll
    for (int i = 0; i < 32; i += 2) { const double2 v = *reinterpret_cast<const double2*>(sm + rep * 32 + i); a32[i] = fma(v.x, x, a32[i]); a32[i + 1] = fma(v.y, x, a32[i + 1]); }
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
  double tot = x + s + r + c;
  for (int i = 0; i < 8; ++i) tot += acc[i];
  for (int i = 0; i < 32; ++i) tot += a32[i];
  out[lane] = tot;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 256); cudaMalloc(&c, 64);
  for (int it = 0; it < 2; ++it) k<<<1, 32>>>(o, c, 1.25);
  long long h[8]; cudaMemcpy(h, c, 56, cudaMemcpyDeviceToHost);
  printf("a) dependent DFMA: %.1f cyc/op\nb) 8 chains DFMA: %.1f cyc/op\nc) indep double shfl + add: %.1f cyc/op\nd) fast_rcp + add chain: %.1f cyc/iter\n"
         "e) bcast LDS.64 + DFMA (8 acc): %.1f cyc/pair\nf) shfl->fma chain: %.1f cyc/iter\ng) LDS.128 + 2 DFMA (32 acc): %.1f cyc per DFMA\n",
         h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 64.0, h[4] / 256.0, h[5] / 64.0, h[6] / 256.0);
  return 0;
}
