// direct.h -- batched DIRECT-L: the derivative-free global search the reference selects for ThompsonSamplingSimple
// (defaultoptions(::Type{<:GPE}, ::Type{ThompsonSamplingSimple}) = (method = :GN_DIRECT_L, restarts = 1, maxeval = 2000),
// src/acquisition.jl:7-9; any acquisition can ask for it through `method`, :24-37 -- the 2nd character 'N' selects the
// derivative-free wrapper, :31-36).
//
// NLopt's GN_DIRECT_L (cdirect.c with which_diam = 1, which_div = 1, which_opt = 1, magic_eps = 0; recalled from the public
// source, not verifiable in this container -- SURVEY App. A) is Gablonsky's locally-biased DIRECT on the unit cube:
//   * size of a hyper-rectangle = half its LONGEST side; at most ONE rectangle per size (the best one, oldest on ties) competes;
//   * the potentially optimal rectangles are the upper-right convex hull of (size, best value) from the largest size down to
//     the size class that holds the incumbent (epsilon = 0);
//   * each is trisected along ALL its longest sides: the 2 n_long centres c +- w/3 e_d are evaluated, the sides are cut in the
//     order of the better of the two values (best first, so the best children keep the largest rectangles).
// What is re-designed: the search is an ask/tell state machine -- ask() returns ALL new centres of one iteration (every
// potentially optimal rectangle, every longest side), the caller evaluates them in ONE device launch (b200bo_acquire_direct,
// capi.cu) and tell()s the values.  The reference evaluates the same points one closure call at a time.  `width` > 1 lets the
// `width` best rectangles of every hull size class divide (a wider batch per launch; 1 = DIRECT-L as described above).
// NLopt's iterates are not reproduced bit for bit (its hull runs on floating-point diameters in a red-black tree; sizes here are
// exact integer trisection levels); oracle/direct_oracle.py restates THIS state machine and the two are compared point for point.
//
// variant 1 = NLopt's GN_DIRECT (Jones' original, cdirect.c with which_diam = 0, which_div = 0, which_opt = 0): size = half the
// DIAGONAL of the rectangle, every rectangle that ties for the best value of a hull size class divides, a cube is trisected along
// all its sides but any other rectangle along ONE longest side only (the first).
//
// Pure host C++ (no CUDA): compiled into libb200bo.so and, for the CPU tests, into oracle/_build/libdirect_host.so.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace b200bo {

struct DirectL {
  static constexpr int MAX_LEVEL = 36;           // 3^-36 ~ 6.7e-18: below this a centre +- w/3 no longer moves in double precision
  int D = 0;
  int64_t maxeval = 0;                           // total evaluations (<= 0: unlimited)
  int width = 1;
  int variant = 0;                               // 0: DIRECT-L (GN_DIRECT_L), 1: DIRECT (GN_DIRECT)
  int64_t evals = 0;
  double best_f = -INFINITY;
  int64_t best_eval = -1;                        // evaluation number (0-based) that produced the incumbent
  std::vector<double> best_c;                    // unit-cube coordinates of the incumbent
  bool finished = false;
  // rectangles, structure of arrays
  std::vector<double> c;                         // [n][D] centres in the unit cube
  std::vector<int8_t> lev;                       // [n][D] trisections per dimension: side_d = 3^-lev_d
  std::vector<double> f;                         // [n] value at the centre (NaN stored as -Inf: never wins)
  std::vector<int> smin;                         // [n] min_d lev_d (the longest side is 3^-smin)
  std::vector<double> size;                      // [n] the size measure: 3^-smin (DIRECT-L) or the diagonal sqrt(sum_d 9^-lev_d) (DIRECT); equal sizes = one class
  double third[MAX_LEVEL + 2];
  // the division in flight
  struct Pending { int rect; int ndim; int dims[64]; };
  std::vector<Pending> pend;
  int64_t pend_points = 0;
  bool started = false;

  void init(int D_, int64_t maxeval_, int width_, int variant_ = 0) {
    D = D_; maxeval = maxeval_; width = width_ < 1 ? 1 : width_; variant = variant_ == 1 ? 1 : 0;
    evals = 0; best_f = -INFINITY; best_eval = -1; best_c.assign(D, 0.5); finished = (D < 1 || D > 64);
    c.clear(); lev.clear(); f.clear(); smin.clear(); size.clear(); pend.clear(); pend_points = 0; started = false;
    third[0] = 1.0;
    for (int k = 1; k < MAX_LEVEL + 2; ++k) third[k] = third[k - 1] / 3.0;      // the same sequence of divisions in the restatement
  }
  int64_t nrect() const { return (int64_t)f.size(); }

  // centres of the next batch (unit cube, point-major [n][D]); returns n, 0 when the search has ended
  int64_t ask(std::vector<double>& pts) {
    pts.clear(); pend.clear(); pend_points = 0;
    if (finished) return 0;
    if (!started) {
      pts.assign(D, 0.5);
      pend_points = 1;
      return 1;
    }
    int64_t room = maxeval > 0 ? maxeval - evals : INT64_MAX;
    if (room <= 0) { finished = true; return 0; }
    // ---- size classes (equal size measure), largest first; per class the dividing candidates, best first (oldest on ties) ----
    std::vector<int> order;
    for (int i = 0; i < (int)f.size(); ++i) if (smin[i] < MAX_LEVEL) order.push_back(i);      // smaller ones can no longer be divided
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return size[x] > size[y]; });
    std::vector<double> cx;                                 // class sizes, descending
    std::vector<std::vector<int>> top;                      // DIRECT-L: the `width` best of the class; DIRECT: every rectangle tied for its best value
    for (size_t o = 0; o < order.size();) {
      size_t e = o;
      while (e < order.size() && size[order[e]] == size[order[o]]) ++e;
      std::vector<int> t;
      if (variant == 0) {
        for (size_t k = o; k < e; ++k) {
          const int i = order[k];
          int pos = (int)t.size();
          while (pos > 0 && f[i] > f[t[pos - 1]]) --pos;    // strict: an older rectangle of equal value stays ahead
          if (pos < width) { t.insert(t.begin() + pos, i); if ((int)t.size() > width) t.pop_back(); }
        }
      } else {
        double fb = -INFINITY;
        for (size_t k = o; k < e; ++k) fb = std::max(fb, f[order[k]]);
        for (size_t k = o; k < e; ++k) if (f[order[k]] == fb) t.push_back(order[k]);          // age order (stable sort)
      }
      cx.push_back(size[order[o]]);
      top.push_back(t);
      o = e;
    }
    // ---- upper-right convex hull over (size, best value), epsilon = 0 ----
    int s_star = -1;                                        // class of the incumbent (largest size on ties)
    for (int s = 0; s < (int)top.size(); ++s)
      if (s_star < 0 || f[top[s][0]] > f[top[s_star][0]]) s_star = s;
    if (s_star < 0) { finished = true; return 0; }
    std::vector<int> hull;                                  // classes, size ascending, starting at s_star
    for (int s = s_star; s >= 0; --s) {
      const double xs = cx[s], ys = f[top[s][0]];
      if (!(ys > -INFINITY) && s != s_star) continue;       // a class that only holds failed evaluations never supports the hull
      while (hull.size() >= 2) {
        const int a = hull[hull.size() - 2], b = hull[hull.size() - 1];
        const double xa = cx[a], ya = f[top[a][0]], xb = cx[b], yb = f[top[b][0]];
        // b is strictly below the chord a -> s  <=>  (yb - ya)(xs - xa) < (ys - ya)(xb - xa)
        if ((yb - ya) * (xs - xa) < (ys - ya) * (xb - xa)) hull.pop_back(); else break;
      }
      hull.push_back(s);
    }
    // ---- the divisions of this iteration, largest rectangles first; the batch is cut at maxeval like NLopt's evaluation counter ----
    for (int hi = (int)hull.size() - 1; hi >= 0 && room > 0; --hi) {
      for (int r : top[hull[hi]]) {
        if (room <= 0) break;
        Pending p;
        p.rect = r; p.ndim = 0;
        const int8_t* lv = lev.data() + (size_t)r * D;
        for (int d = 0; d < D; ++d) if (lv[d] == smin[r]) p.dims[p.ndim++] = d;
        if (variant == 1 && p.ndim < D) p.ndim = 1;         // DIRECT: only a cube is cut along all its sides, anything else along its first longest side
        const double w3 = third[smin[r] + 1];
        const double* cr = c.data() + (size_t)r * D;
        for (int k = 0; k < p.ndim && room > 0; ++k) {
          for (int sgn = 0; sgn < 2 && room > 0; ++sgn) {
            const size_t at = pts.size();
            pts.insert(pts.end(), cr, cr + D);
            pts[at + p.dims[k]] = sgn == 0 ? cr[p.dims[k]] - w3 : cr[p.dims[k]] + w3;
            --room; ++pend_points;
          }
        }
        pend.push_back(p);
      }
    }
    if (pend_points == 0) { finished = true; return 0; }
    return pend_points;
  }

  void note(double v, const double* x) {
    if (v > best_f) { best_f = v; best_eval = evals; best_c.assign(x, x + D); }     // first strict maximum (acquire_max's rule, src/acquisition.jl:62)
    ++evals;
  }
  void push(const double* x, const int8_t* lv, double v) {
    c.insert(c.end(), x, x + D);
    lev.insert(lev.end(), lv, lv + D);
    f.push_back(v);
    int s = lv[0];
    for (int d = 1; d < D; ++d) s = std::min<int>(s, lv[d]);
    smin.push_back(s);
    size.push_back(measure(lv, s));
  }
  // the size measure of a rectangle with trisection levels lv (smallest level s)
  double measure(const int8_t* lv, int s) const {
    if (variant == 0) return third[s];
    int8_t sorted[64];
    std::copy(lv, lv + D, sorted);
    std::sort(sorted, sorted + D);                          // a canonical order: equal level multisets give bit-identical sums
    double q = 0.0;
    for (int d = 0; d < D; ++d) { const double w = third[std::min<int>(sorted[d], MAX_LEVEL + 1)]; q += w * w; }
    return std::sqrt(q);
  }

  // values of the batch handed out by the last ask(), in the same order
  void tell(const std::vector<double>& pts, const double* vals) {
    if (pend_points == 0) return;
    if (!started) {
      started = true;
      const double v = vals[0] == vals[0] ? vals[0] : -INFINITY;
      note(v, pts.data());
      std::vector<int8_t> lv(D, 0);
      push(pts.data(), lv.data(), v);
      if (maxeval > 0 && evals >= maxeval) finished = true;
      return;
    }
    int64_t at = 0;
    for (const Pending& p : pend) {
      // values of this rectangle's 2 ndim points; a division cut short by maxeval only feeds the incumbent
      double fv[2 * 64];
      int have = 0;
      for (int k = 0; k < 2 * p.ndim && at < pend_points; ++k, ++at, ++have) {
        const double v = vals[at];
        fv[k] = v == v ? v : -INFINITY;
        note(fv[k], pts.data() + (size_t)at * D);
      }
      if (have < 2 * p.ndim) { finished = true; break; }
      const int64_t base = at - have;
      int order[64];
      for (int k = 0; k < p.ndim; ++k) order[k] = k;
      std::stable_sort(order, order + p.ndim, [&](int a, int b) { return std::max(fv[2 * a], fv[2 * a + 1]) > std::max(fv[2 * b], fv[2 * b + 1]); });
      std::vector<int8_t> lv(lev.begin() + (size_t)p.rect * D, lev.begin() + (size_t)p.rect * D + D);
      for (int k = 0; k < p.ndim; ++k) {
        const int kk = order[k];
        lv[p.dims[kk]] += 1;
        push(pts.data() + (size_t)(base + 2 * kk) * D, lv.data(), fv[2 * kk]);
        push(pts.data() + (size_t)(base + 2 * kk + 1) * D, lv.data(), fv[2 * kk + 1]);
      }
      std::copy(lv.begin(), lv.end(), lev.begin() + (size_t)p.rect * D);
      int sm = lv[0];
      for (int d = 1; d < D; ++d) sm = std::min<int>(sm, lv[d]);
      smin[p.rect] = sm;
      size[p.rect] = measure(lv.data(), sm);
    }
    pend.clear(); pend_points = 0;
    if (maxeval > 0 && evals >= maxeval) finished = true;
  }
};

}  // namespace b200bo
