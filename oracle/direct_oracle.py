"""TEST INFRASTRUCTURE -- restatement of OUR batched DIRECT-L state machine (bayesianoptimization.jl_b200/csrc/direct.h), which
stands where the reference calls NLopt :GN_DIRECT_L (src/acquisition.jl:7-9, default for ThompsonSamplingSimple: restarts = 1,
maxeval = 2000; derivative-free wrapper :31-36).  NLopt itself is an un-vendored dependency (NLopt.jl 0.4-1 / libnlopt 2.x, no
Manifest): its algorithm -- Gablonsky's locally-biased DIRECT, cdirect.c with longest-side size measure, one rectangle per size,
all longest sides trisected, epsilon = 0 -- is restated from its published description; NLopt's iterates are not reproduced bit for
bit.  Written independently of the C++ (dict/list bookkeeping, a different hull routine); the tests compare the
two point for point.
"""
from __future__ import annotations

import math

import numpy as np

MAX_LEVEL = 36


def _thirds():
    t = [1.0]
    for _ in range(MAX_LEVEL + 1):
        t.append(t[-1] / 3.0)
    return t


def _upper_right_hull(pts):
    """pts: list of (x, y, tag) with x strictly increasing, pts[0] holding the largest y.  Returns the tags on the upper convex hull
    from pts[0] to the right (collinear points kept)."""
    hull = []
    for p in pts:
        while len(hull) >= 2:
            (xa, ya, _), (xb, yb, _) = hull[-2], hull[-1]
            if (yb - ya) * (p[0] - xa) < (p[1] - ya) * (xb - xa):
                hull.pop()
            else:
                break
        hull.append(p)
    return [t for _, _, t in hull]


def _measure(lev, third, variant):
    if variant == 0:
        return third[min(lev)]
    q = 0.0
    for k in sorted(lev):                                             # canonical order: equal level multisets give bit-identical sums
        w = third[min(k, MAX_LEVEL + 1)]
        q += w * w
    return math.sqrt(q)


def direct_l(fbatch, D: int, maxeval: int, width: int = 1, variant: int = 0):
    """fbatch(P) -> values for the rows of P (n x D, unit cube).  variant 0: DIRECT-L (NLopt GN_DIRECT_L), 1: Jones' DIRECT (GN_DIRECT):
    diagonal size measure, every rectangle tied for the best value of a hull class divides, only cubes are cut along all sides.
    Returns dict(best_f, best_c, evals, batches, X (evals x D), f)."""
    third = _thirds()
    rects = []                       # dicts: c (np array), lev (list of int), f
    X_all, f_all = [], []
    best_f, best_c = -math.inf, np.full(D, 0.5)
    evals = 0
    batches = 0

    def clean(v):
        return -math.inf if v != v else float(v)

    # the centre
    P = np.full((1, D), 0.5)
    v = clean(fbatch(P)[0]); batches += 1
    evals += 1; X_all.append(P[0].copy()); f_all.append(v)
    if v > best_f:
        best_f, best_c = v, P[0].copy()
    rects.append(dict(c=P[0].copy(), lev=[0] * D, f=v))
    while evals < maxeval:
        # size classes (equal measure), dividing candidates per class
        classes = {}
        for i, r in enumerate(rects):
            if min(r["lev"]) >= MAX_LEVEL:
                continue
            classes.setdefault(_measure(r["lev"], third, variant), []).append(i)
        if not classes:
            break
        top = {}
        for s, ids in classes.items():
            if variant == 0:
                ids = sorted(ids, key=lambda i: (-rects[i]["f"], i))     # best first, oldest on ties
                top[s] = ids[:width]
            else:
                fb = max(rects[i]["f"] for i in ids)
                top[s] = [i for i in ids if rects[i]["f"] == fb]         # every rectangle tied for the best value, oldest first
        s_star = None
        for s in sorted(top, reverse=True):                              # largest size first: wins ties
            if s_star is None or rects[top[s][0]]["f"] > rects[top[s_star][0]]["f"]:
                s_star = s
        pts = []
        for s in sorted([s for s in top if s >= s_star]):                # size ascending
            y = rects[top[s][0]]["f"]
            if y == -math.inf and s != s_star:
                continue
            pts.append((s, y, s))
        hull = _upper_right_hull(pts)
        # divisions, largest rectangles first, cut at maxeval
        room = maxeval - evals
        plan, newp = [], []
        for s in reversed(hull):
            for i in top[s]:
                if room <= 0:
                    break
                r = rects[i]
                smin = min(r["lev"])
                dims = [d for d in range(D) if r["lev"][d] == smin]
                if variant == 1 and len(dims) < D:
                    dims = dims[:1]                                      # only a cube is cut along all its sides
                w3 = third[smin + 1]
                cnt = 0
                for d in dims:
                    for sgn in (-1.0, 1.0):
                        if room <= 0:
                            break
                        x = r["c"].copy()
                        x[d] = r["c"][d] - w3 if sgn < 0 else r["c"][d] + w3
                        newp.append(x); room -= 1; cnt += 1
                plan.append((i, dims, cnt))
        if not newp:
            break
        P = np.array(newp)
        vals = [clean(v) for v in fbatch(P)]; batches += 1
        at = 0
        for i, dims, cnt in plan:
            fv = vals[at:at + cnt]
            for k in range(cnt):
                evals += 1; X_all.append(P[at + k].copy()); f_all.append(fv[k])
                if fv[k] > best_f:
                    best_f, best_c = fv[k], P[at + k].copy()
            if cnt < 2 * len(dims):
                at += cnt
                break
            order = sorted(range(len(dims)), key=lambda k: -max(fv[2 * k], fv[2 * k + 1]))   # stable: ties keep dimension order
            lev = list(rects[i]["lev"])
            for k in order:
                lev[dims[k]] += 1
                rects.append(dict(c=P[at + 2 * k].copy(), lev=list(lev), f=fv[2 * k]))
                rects.append(dict(c=P[at + 2 * k + 1].copy(), lev=list(lev), f=fv[2 * k + 1]))
            rects[i]["lev"] = lev
            at += cnt
    return dict(best_f=best_f, best_c=best_c, evals=evals, batches=batches, X=np.array(X_all), f=np.array(f_all), nrect=len(rects))
