// acqfn.cuh -- the acquisition functors as coded in the reference, the Philox stream of ThompsonSamplingSimple and the selection rule
// of acquire_max, shared by the two engines of K6 (acq.cu: DMMA solve, acq_i8.cu: tcgen05 int8-slice product).
#pragma once
#include "common.cuh"
#include "../../include/b200bo.h"

namespace b200bo {

// ---- Philox4x32-10 keyed by (seed, global candidate index); identical to oracle/gp_oracle.py:philox_normal ----
__device__ __forceinline__ double philox_normal(unsigned long long seed, unsigned long long idx) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = 0u, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const unsigned long long a = ((unsigned long long)c0 << 32) | c1, b = ((unsigned long long)c2 << 32) | c3;
  const double u1 = ((double)(a >> 11) + 1.0) * 1.1102230246251565e-16;
  const double u2 = (double)(b >> 11) * 1.1102230246251565e-16;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

// ---- the functors AS CODED in the reference (quirks 1-3 of SURVEY 0.4) and their partials (App. A table) ----
__device__ __forceinline__ void acq_eval(int kind, double p0, double p1, double mu, double s2, double eps, double& val, double& amu,
                                         double& as2) {
  const double INV_SQRT_2PI = 0.3989422804014327;
  amu = 1.0; as2 = 0.0;
  switch (kind) {
    case B200BO_ACQ_PI:
    case B200BO_ACQ_EI: {
      const double d = mu - p0;
      if (s2 == 0.0) {
        const double gt = mu > p0 ? 1.0 : 0.0;
        val = (kind == B200BO_ACQ_PI) ? gt : (mu > p0 ? d : 0.0);
        amu = (kind == B200BO_ACQ_PI) ? 0.0 : gt;
        return;
      }
      const double sig = sqrt(s2);
      const double cdf = 0.5 * (1.0 + erf(d / sqrt(2.0 * s2)));                            // utils.jl:49
      const double z = d / sig;
      const double ph = INV_SQRT_2PI * exp(-0.5 * z * z);
      if (kind == B200BO_ACQ_PI) {
        val = cdf; amu = ph / sig; as2 = -z * ph / (2.0 * s2);
      } else {
        const double pdf = 1.0 / sqrt(6.283185307179586 * s2) * exp(-(d * d) / (2.0 * s2));  // utils.jl:48
        val = d * cdf + sig * pdf;                                                           // acquisitionfunctions.jl:49
        amu = cdf + z * ph * (1.0 - 1.0 / sig);
        as2 = z * z * (1.0 - sig) * ph / (2.0 * s2);
      }
      return;
    }
    case B200BO_ACQ_UCB: {
      const double sig = sqrt(s2);
      val = mu + p0 * sig;
      as2 = s2 == 0.0 ? 0.0 : p0 / (2.0 * sig);
      return;
    }
    case B200BO_ACQ_MI: {
      const double den = sqrt(s2 + p1);
      val = mu + p0 * (den - sqrt(p1));
      as2 = den == 0.0 ? 0.0 : p0 / (2.0 * den);
      return;
    }
    case B200BO_ACQ_TS:
      val = mu + sqrt(s2) * eps;
      return;
    default:
      val = mu;
      return;
  }
}

__device__ __forceinline__ bool better(double v, int64_t i, double bv, int64_t bi) {
  return (v > bv) || (v == bv && bi >= 0 && i < bi);
}

}  // namespace b200bo
