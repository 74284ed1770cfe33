"""TEST INFRASTRUCTURE ONLY: plain-Python restatement of bayesianoptimization.jl_b200/csrc/lbfgs.cuh (the box-bounded L-BFGS ascent that
stands in for NLopt's :LD_LBFGS runs of the reference, src/acquisition.jl:24-37,59 and src/models/gp.jl:65-74): same state machine, same
constants, same stopping rule.  NLopt's own iterates are not reproduced; what is pinned here is OUR replacement."""
import numpy as np

LB_M = 6
FTOL, XTOL, MAXEVAL, STALLED = 3, 4, 5, 6


class Run:
    def __init__(self, x0, lb, ub, maxeval=0, ftol_rel=0.0, ftol_abs=0.0, xtol_rel=0.0, xtol_abs=0.0, step0=0.02):
        self.lb, self.ub = np.asarray(lb, float), np.asarray(ub, float)
        self.o = dict(maxeval=maxeval, ftol_rel=ftol_rel, ftol_abs=ftol_abs, xtol_rel=xtol_rel, xtol_abs=xtol_abs, step0=step0)
        self.xe = np.clip(np.asarray(x0, float), self.lb, self.ub)
        D = self.xe.size
        self.x = np.zeros(D); self.g = np.zeros(D); self.d = np.zeros(D)
        self.S = np.zeros((LB_M, D)); self.Y = np.zeros((LB_M, D)); self.rho = np.zeros(LB_M)
        self.f = 0.0; self.t = 0.0; self.nhist = 0; self.head = 0; self.phase = 0; self.evals = 0; self.status = 0

    def _direction(self, first):
        nh, head = self.nhist, self.head
        a = np.zeros(LB_M)
        at_lo, at_hi = self.x <= self.lb, self.x >= self.ub
        free = ~((at_lo & (self.g < 0)) | (at_hi & (self.g > 0)))          # the recursion runs in the subspace of the free coordinates
        self.d = np.where(free, self.g, 0.0)
        for j in range(nh):
            p = (head - 1 - j) % LB_M
            a[j] = self.rho[p] * (self.S[p] @ self.d)
            self.d = np.where(free, self.d - a[j] * self.Y[p], 0.0)
        if nh > 0:
            p = (head - 1) % LB_M
            yy = np.sum(np.where(free, self.Y[p] * self.Y[p], 0.0))
            sy = np.sum(np.where(free, self.S[p] * self.Y[p], 0.0))
            self.d = self.d * (sy / yy if (yy > 0 and sy > 0) else 1.0)
        for j in range(nh - 1, -1, -1):
            p = (head - 1 - j) % LB_M
            b = self.rho[p] * (self.Y[p] @ self.d)
            self.d = np.where(free, self.d + (a[j] - b) * self.S[p], 0.0)
        fix = (at_lo & ((self.d < 0) | (self.g < 0))) | (at_hi & ((self.d > 0) | (self.g > 0)))
        self.d = np.where(fix, 0.0, self.d)
        gd, dn = self.g @ self.d, self.d @ self.d
        w = self.ub - self.lb
        diag = float(np.sum(np.where(np.isfinite(w), w * w, 0.0)))
        if not (gd > 0.0) or not (dn < np.inf):
            self.nhist = 0; self.head = 0
            self.d = np.where((at_lo & (self.g < 0)) | (at_hi & (self.g > 0)), 0.0, self.g)
            gd, dn = self.g @ self.d, self.d @ self.d
            first = True
        if not (dn > 0.0) or gd != gd:
            return False
        t = 1.0
        if first:
            t = self.o["step0"] * (np.sqrt(diag) if diag > 0 else 1.0) / np.sqrt(dn)
        self.t = t
        self.xe = np.clip(self.x + t * self.d, self.lb, self.ub)
        return bool(np.any(self.xe != self.x))

    def step(self, f_new, g_new):
        """consume the evaluation of self.xe; afterwards self.xe is the next point (or the result when self.status != 0)"""
        if self.status:
            self.xe = self.x.copy(); return
        o = self.o
        self.evals += 1
        out = o["maxeval"] > 0 and self.evals >= o["maxeval"]
        stop, need_dir = 0, False
        g_new = np.asarray(g_new, float)
        if self.phase == 0:
            self.x, self.g, self.f = self.xe.copy(), g_new.copy(), float(f_new)
            self.nhist = self.head = 0; self.phase = 1
            if f_new != f_new: stop = STALLED
            elif out: stop = MAXEVAL
            else: need_dir = True
        else:
            lin = self.g @ (self.xe - self.x)
            if f_new >= self.f + 1e-4 * lin and f_new == f_new:
                s, yv = self.xe - self.x, self.g - g_new
                sy, ss, yy = s @ yv, s @ s, yv @ yv
                xsmall = bool(np.all((np.abs(s) <= o["xtol_rel"] * np.abs(self.xe)) | (np.abs(s) <= o["xtol_abs"]))) and (o["xtol_rel"] > 0 or o["xtol_abs"] > 0)
                df = abs(f_new - self.f)
                fsmall = (o["ftol_rel"] > 0 and df <= o["ftol_rel"] * abs(f_new)) or (o["ftol_abs"] > 0 and df <= o["ftol_abs"])
                if sy > 1e-10 * np.sqrt(ss * yy) and sy > 0:
                    p = self.head % LB_M
                    self.S[p], self.Y[p], self.rho[p] = s, yv, 1.0 / sy
                    self.head = (p + 1) % LB_M
                    self.nhist = min(self.nhist + 1, LB_M)
                self.x, self.g, self.f = self.xe.copy(), g_new.copy(), float(f_new)
                if fsmall: stop = FTOL
                elif xsmall: stop = XTOL
                elif out: stop = MAXEVAL
                else: need_dir = True
            elif out:
                stop = MAXEVAL
            else:
                curv = f_new - self.f - lin                       # quadratic model of f along the path: f + lin tau + curv tau^2, tau in [0, 1]
                tau = -lin / (2.0 * curv) if (curv < 0.0 and f_new == f_new) else 0.5
                self.t *= min(0.5, max(0.1, tau))
                self.xe = np.clip(self.x + self.t * self.d, self.lb, self.ub)
                if not np.any(np.abs(self.xe - self.x) > 1e-14 * (np.abs(self.x) + 1e-300)):
                    stop = XTOL
        if need_dir and not self._direction(self.nhist == 0):
            stop = XTOL
        if stop:
            self.status = stop; self.xe = self.x.copy()


def maximize(fg, x0, lb, ub, **opts):
    """fg(x) -> (value, gradient).  Returns the finished Run (x, f, evals, status)."""
    r = Run(x0, lb, ub, **opts)
    cap = opts.get("maxeval", 0) or 100000
    for _ in range(cap):
        f, g = fg(r.xe)
        r.step(f, g)
        if r.status:
            break
    return r
