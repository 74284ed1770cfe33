// TEST INFRASTRUCTURE -- host-only C entry points around the product's DIRECT-L state machine (csrc/direct.h is pure C++), so the CPU
// suite can compare it point for point with oracle/direct_oracle.py without a GPU.  Built into oracle/_build/libdirect_host.so.
#include "../bayesianoptimization.jl_b200/csrc/direct.h"
#include <cstring>

extern "C" {
void* direct_host_create(int D, long long maxeval, int width, int variant) {
  auto* s = new b200bo::DirectL();
  s->init(D, maxeval, width, variant);
  return s;
}
void direct_host_destroy(void* p) { delete static_cast<b200bo::DirectL*>(p); }
static thread_local std::vector<double> g_pts;
// returns the number of points of the next batch and copies them (n x D, row = point) into out (capacity cap points)
long long direct_host_ask(void* p, double* out, long long cap) {
  auto* s = static_cast<b200bo::DirectL*>(p);
  const long long n = s->ask(g_pts);
  if (n > cap) return -n;
  if (n > 0) memcpy(out, g_pts.data(), sizeof(double) * n * s->D);
  return n;
}
void direct_host_tell(void* p, const double* vals) {
  auto* s = static_cast<b200bo::DirectL*>(p);
  s->tell(g_pts, vals);
}
void direct_host_result(void* p, double* best_f, double* best_c, long long* evals, long long* nrect) {
  auto* s = static_cast<b200bo::DirectL*>(p);
  *best_f = s->best_f; *evals = s->evals; *nrect = s->nrect();
  memcpy(best_c, s->best_c.data(), sizeof(double) * s->D);
}
}
