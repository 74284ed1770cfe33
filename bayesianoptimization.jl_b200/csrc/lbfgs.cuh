// lbfgs.cuh -- one step of a box-bounded limited-memory BFGS ASCENT, written once for the device (one thread per restart, M restarts in
// lock-step on the fused value + gradient launch) and for the host (the MAP fit over K7, a handful of hyper-parameters).
//
// What it stands in for: NLopt's :LD_LBFGS runs of the reference -- one per restart in acquire_max (src/acquisition.jl:59, bounds
// :28-29) and one per MAP fit in optimizemodel! (src/models/gp.jl:69-74, bounds :65-68) -- with the option set the reference forwards
// (src/acquisition.jl:24-27): maxeval counts objective evaluations per run; ftol_rel / ftol_abs stop on the change of the objective
// between two accepted iterates, xtol_rel / xtol_abs on the change of every coordinate; 0 disables a criterion, as in NLopt.
// NLopt's own iterates (Luksan's PLIS) are not reproduced -- SURVEY App. A -- the contract is the stopping rule, the bounds and a
// monotone sequence of accepted iterates.
//
// Method: two-loop recursion on the last `m` curvature pairs, search direction projected on the active set of the box, backtracking on
// the projected path  x(t) = P(x + t d)  with the sufficient-increase test  f(x(t)) >= f(x) + c1 g'(x(t) - x).
// One call consumes ONE evaluation (value f_new, gradient g_new at the point the previous call left in `xe`) and leaves the next
// point to evaluate in `xe`; `status` != 0 means the run has stopped and x / f hold its result.
#pragma once
#include <math.h>
#include <stdint.h>

namespace b200bo {

struct LbfgsOpts {
  int maxeval;                 // <= 0: unlimited
  double ftol_rel, ftol_abs, xtol_rel, xtol_abs;
  double step0;                // length of the very first trial step, as a fraction of the box diagonal
};

enum { LB_RUNNING = 0, LB_FTOL = 3, LB_XTOL = 4, LB_MAXEVAL = 5, LB_STALLED = 6 };   // NLopt's FTOL_REACHED / XTOL_REACHED / MAXEVAL_REACHED numbering
constexpr int LB_M = 6;        // curvature pairs kept

// state of one run: [x D][g D][d D][S m D][Y m D][rho m][f, t, nhist, head, phase, evals, status, spare]
__host__ __device__ inline int64_t lbfgs_state_doubles(int D) { return 3 * (int64_t)D + 2 * (int64_t)LB_M * D + LB_M + 8; }

struct LbfgsState {
  double *x, *g, *d, *S, *Y, *rho, *sc;
  __host__ __device__ LbfgsState(double* st, int D) : x(st), g(st + D), d(st + 2 * D), S(st + 3 * D), Y(st + 3 * D + LB_M * D), rho(st + 3 * D + 2 * LB_M * D), sc(st + 3 * D + 2 * LB_M * D + LB_M) {}
  __host__ __device__ double& f() { return sc[0]; }
  __host__ __device__ double& t() { return sc[1]; }
  __host__ __device__ double& nhist() { return sc[2]; }
  __host__ __device__ double& head() { return sc[3]; }
  __host__ __device__ double& phase() { return sc[4]; }
  __host__ __device__ double& evals() { return sc[5]; }
  __host__ __device__ double& status() { return sc[6]; }
};

__host__ __device__ inline double lb_clamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// new search direction at (x, g) and the first trial point of its line search; returns false when the projected step does not move
__host__ __device__ inline bool lbfgs_direction(LbfgsState& s, double* xe, const double* lb, const double* ub, int D, const LbfgsOpts& o, bool first) {
  const int nh = (int)s.nhist(), head = (int)s.head();
  double a[LB_M];
  // the recursion runs in the subspace of the FREE coordinates (not on a bound with the gradient pointing outward): reduced quasi-Newton step
  unsigned long long freem = 0ull;
  for (int k = 0; k < D; ++k) {
    const bool at_lo = s.x[k] <= lb[k], at_hi = s.x[k] >= ub[k];
    const bool fr = !((at_lo && s.g[k] < 0.0) || (at_hi && s.g[k] > 0.0));
    if (fr) freem |= 1ull << k;
    s.d[k] = fr ? s.g[k] : 0.0;
  }
#define LB_FREE(k) ((freem >> (k)) & 1ull)
  for (int j = 0; j < nh; ++j) {                       // newest -> oldest
    const int p = ((head - 1 - j) % LB_M + LB_M) % LB_M;
    double sq = 0.0;
    for (int k = 0; k < D; ++k) sq = fma(s.S[p * D + k], s.d[k], sq);
    a[j] = s.rho[p] * sq;
    for (int k = 0; k < D; ++k) if (LB_FREE(k)) s.d[k] = fma(-a[j], s.Y[p * D + k], s.d[k]);
  }
  if (nh > 0) {
    const int p = ((head - 1) % LB_M + LB_M) % LB_M;
    double yy = 0.0, sy = 0.0;
    for (int k = 0; k < D; ++k) if (LB_FREE(k)) { yy = fma(s.Y[p * D + k], s.Y[p * D + k], yy); sy = fma(s.S[p * D + k], s.Y[p * D + k], sy); }
    const double gamma = (yy > 0.0 && sy > 0.0) ? sy / yy : 1.0;
    for (int k = 0; k < D; ++k) s.d[k] *= gamma;
  }
  for (int j = nh - 1; j >= 0; --j) {                  // oldest -> newest
    const int p = ((head - 1 - j) % LB_M + LB_M) % LB_M;
    double yr = 0.0;
    for (int k = 0; k < D; ++k) yr = fma(s.Y[p * D + k], s.d[k], yr);
    const double b = s.rho[p] * yr;
    for (int k = 0; k < D; ++k) if (LB_FREE(k)) s.d[k] = fma(a[j] - b, s.S[p * D + k], s.d[k]);
  }
#undef LB_FREE
  // active set of the box: a coordinate on a bound whose direction (or gradient) points outward stays there
  double gd = 0.0, dn = 0.0, diag = 0.0;
  for (int k = 0; k < D; ++k) {
    const bool at_lo = s.x[k] <= lb[k], at_hi = s.x[k] >= ub[k];
    if ((at_lo && (s.d[k] < 0.0 || s.g[k] < 0.0)) || (at_hi && (s.d[k] > 0.0 || s.g[k] > 0.0))) s.d[k] = 0.0;
    gd = fma(s.g[k], s.d[k], gd);
    dn = fma(s.d[k], s.d[k], dn);
    const double w = ub[k] - lb[k];
    if (w < INFINITY) diag = fma(w, w, diag);
  }
  if (!(gd > 0.0) || !(dn < INFINITY)) {               // not an ascent direction (or non-finite): steepest ascent on the free set, forget the pairs
    s.nhist() = 0.0; s.head() = 0.0;
    gd = 0.0; dn = 0.0;
    for (int k = 0; k < D; ++k) {
      const bool at_lo = s.x[k] <= lb[k], at_hi = s.x[k] >= ub[k];
      s.d[k] = ((at_lo && s.g[k] < 0.0) || (at_hi && s.g[k] > 0.0)) ? 0.0 : s.g[k];
      gd = fma(s.g[k], s.d[k], gd);
      dn = fma(s.d[k], s.d[k], dn);
    }
    first = true;
  }
  if (!(dn > 0.0) || !(gd == gd)) return false;        // projected gradient is zero (or NaN): stationary
  double t = 1.0;
  if (first) {                                         // no curvature information: a step of fixed length
    const double len = o.step0 * (diag > 0.0 ? sqrt(diag) : 1.0);
    t = len / sqrt(dn);
  }
  s.t() = t;
  bool moved = false;
  for (int k = 0; k < D; ++k) { xe[k] = lb_clamp(fma(t, s.d[k], s.x[k]), lb[k], ub[k]); moved = moved || xe[k] != s.x[k]; }
  return moved;
}

// consume the evaluation (f_new, g_new) of the point in xe; leave the next point in xe (or the result, when stopped)
__host__ __device__ inline void lbfgs_step(double* st, double* xe, double f_new, const double* g_new, const double* lb, const double* ub, int D,
                                           const LbfgsOpts& o) {
  LbfgsState s(st, D);
  if (s.status() != 0.0) { for (int k = 0; k < D; ++k) xe[k] = s.x[k]; return; }
  const double c1 = 1e-4;
  s.evals() += 1.0;
  const bool out_of_evals = o.maxeval > 0 && s.evals() >= (double)o.maxeval;
  int stop = 0;
  bool need_dir = false;
  if (s.phase() == 0.0) {                              // the start itself
    for (int k = 0; k < D; ++k) { s.x[k] = xe[k]; s.g[k] = g_new[k]; }
    s.f() = f_new; s.nhist() = 0.0; s.head() = 0.0; s.phase() = 1.0;
    if (!(f_new == f_new)) stop = LB_STALLED;          // NaN objective at the start: nothing to climb
    else if (out_of_evals) stop = LB_MAXEVAL;
    else need_dir = true;
  } else {
    double lin = 0.0;
    for (int k = 0; k < D; ++k) lin = fma(s.g[k], xe[k] - s.x[k], lin);
    if (f_new >= s.f() + c1 * lin && f_new == f_new) { // sufficient increase: accept the iterate
      double sy = 0.0, ss = 0.0, yy = 0.0;
      bool xsmall = true;
      for (int k = 0; k < D; ++k) {
        const double sk = xe[k] - s.x[k], yk = s.g[k] - g_new[k];     // curvature pair of -f
        sy = fma(sk, yk, sy); ss = fma(sk, sk, ss); yy = fma(yk, yk, yy);
        if (!(fabs(sk) <= o.xtol_rel * fabs(xe[k]) || fabs(sk) <= o.xtol_abs)) xsmall = false;
      }
      const double df = fabs(f_new - s.f());
      const bool fsmall = (o.ftol_rel > 0.0 && df <= o.ftol_rel * fabs(f_new)) || (o.ftol_abs > 0.0 && df <= o.ftol_abs);
      xsmall = xsmall && (o.xtol_rel > 0.0 || o.xtol_abs > 0.0);
      if (sy > 1e-10 * sqrt(ss * yy) && sy > 0.0) {    // keep the pair only if it carries positive curvature (it replaces the oldest)
        const int p = (int)s.head() % LB_M;
        for (int k = 0; k < D; ++k) { s.S[p * D + k] = xe[k] - s.x[k]; s.Y[p * D + k] = s.g[k] - g_new[k]; }
        s.rho[p] = 1.0 / sy;
        s.head() = (double)((p + 1) % LB_M);
        if (s.nhist() < (double)LB_M) s.nhist() += 1.0;
      }
      for (int k = 0; k < D; ++k) { s.x[k] = xe[k]; s.g[k] = g_new[k]; }
      s.f() = f_new;
      if (fsmall) stop = LB_FTOL;
      else if (xsmall) stop = LB_XTOL;
      else if (out_of_evals) stop = LB_MAXEVAL;
      else need_dir = true;
    } else {                                           // backtrack on the projected path
      if (out_of_evals) stop = LB_MAXEVAL;
      else {
        const double curv = f_new - s.f() - lin;        // quadratic model of f along the path: f + lin tau + curv tau^2, tau in [0, 1]
        const double tau = (curv < 0.0 && f_new == f_new) ? -lin / (2.0 * curv) : 0.5;
        s.t() *= fmin(0.5, fmax(0.1, tau));
        bool moved = false;
        for (int k = 0; k < D; ++k) { xe[k] = lb_clamp(fma(s.t(), s.d[k], s.x[k]), lb[k], ub[k]); moved = moved || xe[k] != s.x[k]; }
        bool tiny = true;
        for (int k = 0; k < D; ++k) if (fabs(xe[k] - s.x[k]) > 1e-14 * (fabs(s.x[k]) + 1e-300)) tiny = false;
        if (!moved || tiny) stop = LB_XTOL;            // the step has shrunk below the resolution of x
      }
    }
  }
  if (need_dir && !lbfgs_direction(s, xe, lb, ub, D, o, s.nhist() == 0.0)) stop = LB_XTOL;
  if (stop) { s.status() = (double)stop; for (int k = 0; k < D; ++k) xe[k] = s.x[k]; }
}

}  // namespace b200bo
