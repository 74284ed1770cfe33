// common.cuh -- shared device helpers for libb200bo (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace b200bo {

constexpr int NB = 128;        // block size of the blocked factor (rows/cols per panel)
constexpr int KC = 16;         // k-chunk in doubles per pipeline stage (128 B per row)
constexpr int TILE_N = 64;     // candidate columns per acquisition tile

// ---- kernel families: k = sf2 * phi(r2), psi = -2 dphi/dr2 (oracle/gp_oracle.py:_phi_psi) ----
enum { FAM_SE = 0, FAM_MAT12 = 1, FAM_MAT32 = 2, FAM_MAT52 = 3 };

// exp(x) for x <= 0 (or NaN), branch-free and table-free on the FP64 FMA pipe: n = rint(x log2 e) by the 1.5*2^52
// trick, r = x - n ln2 (two-term Cody-Waite), degree-13 Taylor polynomial on |r| <= ln2/2 (truncation 4e-18), 2^n added
// into the exponent field.  ~17 FP64 + 4 integer instructions; <= 1 ulp from libdevice exp().  x < -708 -> 0.
__device__ __forceinline__ double exp_neg(double x) {
  const double MAGIC = 6755399441055744.0;
  const double t = fma(x, 1.4426950408889634, MAGIC);
  const int n = __double2loint(t);
  const double fn = t - MAGIC;
  double r = fma(fn, -6.93147180369123816490e-01, x);
  r = fma(fn, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;            // 1/13!
  p = fma(p, r, 2.08767569878681e-09);          // 1/12!
  p = fma(p, r, 2.505210838544172e-08);         // 1/11!
  p = fma(p, r, 2.755731922398589e-07);         // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
  p = fma(p, r, 2.48015873015873e-05);          // 1/8!
  p = fma(p, r, 1.984126984126984e-04);         // 1/7!
  p = fma(p, r, 1.388888888888889e-03);         // 1/6!
  p = fma(p, r, 8.333333333333333e-03);         // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  res = x < -708.0 ? 0.0 : res;
  return x != x ? x : res;
}

// exp(x) * T[0] for x <= 0 with a 16-entry table T[j] = c * 2^(j/16) (shared memory; c folds a constant factor in):
// n = rint(16 x log2 e), x = n ln2/16 + r, |r| <= ln2/32, degree-7 Taylor polynomial (truncation 1.2e-18), 2^(n>>4) added into
// the exponent field.  12 FP64 instructions (exp_neg: 17).  x < -670 -> 0 (keeps c in [2^-40, 2^40] clear of denormals).
__device__ __forceinline__ double exp_neg16(double x, const double* __restrict__ T) {
  const double MAGIC = 6755399441055744.0;
  const double t = fma(x, 23.083120654223414, MAGIC);           // 16 / ln 2
  const int n = __double2loint(t);
  const double fn = t - MAGIC;
  double r = fma(fn, -4.33216987730702385306e-02, x);            // ln2_hi / 16
  r = fma(fn, -1.19263433079411731251e-11, r);                   // ln2_lo / 16
  double p = 1.984126984126984e-04;             // 1/7!
  p = fma(p, r, 1.388888888888889e-03);         // 1/6!
  p = fma(p, r, 8.333333333333333e-03);         // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= T[n & 15];
  double res = __hiloint2double(__double2hiint(p) + ((n >> 4) << 20), __double2loint(p));
  res = x < -670.0 ? 0.0 : res;
  return x != x ? x : res;
}

template <int FAM>
__device__ __forceinline__ double kern_phi(double r2) {
  if (FAM == FAM_SE) return exp_neg(-0.5 * r2);
  const double r = sqrt(r2);
  if (FAM == FAM_MAT12) return exp_neg(-r);
  if (FAM == FAM_MAT32) { const double s = 1.7320508075688772 * r; return (1.0 + s) * exp_neg(-s); }
  const double s = 2.23606797749979 * r;
  return (1.0 + s + s * s * (1.0 / 3.0)) * exp_neg(-s);
}

template <int FAM>
__device__ __forceinline__ void kern_phi_psi(double r2, double& phi, double& psi) {
  if (FAM == FAM_SE) { phi = exp_neg(-0.5 * r2); psi = phi; return; }
  const double r = sqrt(r2);
  if (FAM == FAM_MAT12) { phi = exp_neg(-r); psi = r > 0.0 ? phi / r : 0.0; return; }
  if (FAM == FAM_MAT32) { const double s = 1.7320508075688772 * r; const double e = exp_neg(-s); phi = (1.0 + s) * e; psi = 3.0 * e; return; }
  const double s = 2.23606797749979 * r; const double e = exp_neg(-s);
  phi = (1.0 + s + s * s * (1.0 / 3.0)) * e; psi = (5.0 / 3.0) * (1.0 + s) * e;
}

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  lane = 4*g + q:
//      a = A[g][q], b = B[q][g], c0/c1 = C[g][2q], C[g][2q+1]  (SASS: DMMA.8x8x4) ----
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

}  // namespace b200bo
