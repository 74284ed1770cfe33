"""Acquisition search (src/acquisition.jl): the multi-restart maximisation, re-designed as ONE batched launch.

The reference runs `restarts` serial NLopt optimisations, each calling the closure up to `maxeval` times
(acquisition.jl:54-68).  Here `restarts` becomes the number M of Latin-hypercube candidates scored -- value,
optional gradient and arg-max -- by a single fused kernel launch (b200bo_acquire).  `polish` then refines the `polish_top`
best candidates by box-bounded L-BFGS ascents in lock-step INSIDE the library (b200bo_acquire_lbfgs) -- what the reference's
NLopt LD_LBFGS run does per start -- with the options it forwards (maxeval, maxtime, ftol_rel/abs, xtol_rel/abs).
With method = :GN_DIRECT_L (the reference's default for ThompsonSamplingSimple) the library's batched DIRECT-L search
(b200bo_acquire_direct, `maxeval` evaluations) runs besides the sweep and the better of the two is returned.
"""
from __future__ import annotations

import numpy as np

from .acquisitionfunctions import (AbstractAcquisition, MaxMean, ThompsonSamplingSimple, setparams)
from .gp import B200GPE
from .utils import ScaledLHSIterator


def defaultoptions(model_type, acq_type) -> dict:
    """defaultoptions(::Type{<:Model}, ::Type{<:AbstractAcquisition}) (acquisition.jl:4-9).  The B200 model type
    keeps the reference's keys; `restarts` defaults to a batch that fills the chip."""
    if acq_type is ThompsonSamplingSimple or (isinstance(acq_type, type) and issubclass(acq_type, ThompsonSamplingSimple)):
        return dict(method="GN_DIRECT_L", restarts=16384, maxeval=2000)
    return dict(method="LD_LBFGS", restarts=16384, maxeval=2000)


class AcquisitionSearch:
    """Stands where the reference keeps an `NLopt.Opt` (BOpt.opt, BayesianOptimization.jl:74,134): options are
    attributes (test/acquisition.jl:7-9 reads opt.maxeval, opt.maxtime, opt.ftol_abs)."""

    def __init__(self, a: AbstractAcquisition, model: B200GPE, lowerbounds, upperbounds, options: dict):
        self.acquisition, self.model = a, model
        self.lower_bounds = np.asarray(lowerbounds, float).copy()
        self.upper_bounds = np.asarray(upperbounds, float).copy()
        self.method = options.get("method", "LD_LBFGS")
        self.maxeval, self.maxtime = 0, 0.0
        self.ftol_abs = self.ftol_rel = self.xtol_abs = self.xtol_rel = 0.0
        self.polish = True
        self.polish_top = 16             # L-BFGS restarts refined in lock-step after the sweep (the reference refines every start)
        self.device_lhs = False          # candidates generated in HBM (b200bo_acquire_lhs) instead of on the host
        self.ascent_steps = 0            # > 0: refine the `ascent_top` best candidates by batched ascent on the device
        self.ascent_top = 256
        self.direct = True               # method :GN_DIRECT*: run the batched DIRECT-L search (b200bo_acquire_direct) besides the sweep
        self.direct_width = 1            # rectangles divided per size class and iteration (1 = DIRECT-L)
        self.seed = 0
        self.rng = None
        for k, v in options.items():                 # acquisition.jl:24-27: every key but method/restarts is set
            if k in ("method", "restarts"):
                continue
            setattr(self, k, v)
        self.gradient = str(self.method)[1] == "D"   # acquisition.jl:31: 2nd character of the method name
        self.last = None


def nlopt_setup(a: AbstractAcquisition, model: B200GPE, lowerbounds, upperbounds, options: dict) -> AcquisitionSearch:
    """acquisition.jl:20-38."""
    opt = AcquisitionSearch(a, model, lowerbounds, upperbounds, options)
    setparams(a, model)                              # :30
    return opt


def _refine(opt: AcquisitionSearch, data: np.ndarray, values: np.ndarray):
    """the NLopt :LD_LBFGS runs of the reference (acquisition.jl:59), for the `polish_top` best candidates of the sweep at once: box-bounded
    L-BFGS ascents in lock-step inside the library (b200bo_acquire_lbfgs) with the forwarded options (acquisition.jl:24-27)."""
    a, model = opt.acquisition, opt.model
    v = np.where(np.isnan(values), -np.inf, values)
    top = np.argsort(-v, kind="stable")[:max(1, min(int(opt.polish_top), v.size))]
    return model.acquire_lbfgs(a.kind, a.params(), data[:, top], opt.lower_bounds, opt.upper_bounds, maxeval=int(opt.maxeval) if opt.maxeval else 2000,
                               ftol_rel=float(opt.ftol_rel), ftol_abs=float(opt.ftol_abs), xtol_rel=float(opt.xtol_rel), xtol_abs=float(opt.xtol_abs),
                               maxtime=float(opt.maxtime))


def acquire_max(opt, lowerbounds=None, upperbounds=None, restarts=None, options=None):
    """acquire_max (acquisition.jl:42-68).  Forms:
         acquire_max(opt, lb, ub, restarts)              (:54)   opt from nlopt_setup
         acquire_max(a, model, lb, ub, options)          (:49)
       Returns (maxf, maxx): first strict maximum over the candidates; if nothing beats -Inf, (-Inf, lowerbounds)."""
    if isinstance(opt, AbstractAcquisition):         # (a, model, lb, ub, options)
        a, model, lb, ub, options = opt, lowerbounds, upperbounds, restarts, options
        o = nlopt_setup(a, model, lb, ub, options)
        return acquire_max(o, lb, ub, options["restarts"])
    lb = np.asarray(lowerbounds, float); ub = np.asarray(upperbounds, float)
    a, model = opt.acquisition, opt.model
    derivative_free = isinstance(a, ThompsonSamplingSimple) or not opt.gradient
    refine = (opt.polish or opt.ascent_steps > 0) and not derivative_free
    if opt.device_lhs and not refine:
        r = model.acquire_lhs(a.kind, a.params(), lb, ub, int(restarts), lhs_seed=opt.seed, ts_seed=opt.seed)
    else:
        data = model.lhs(lb, ub, int(restarts), seed=opt.seed) if opt.device_lhs else ScaledLHSIterator(lb, ub, int(restarts), opt.rng).data  # :57
        r = model.acquire(a.kind, a.params(), data, seed=opt.seed, want_values=refine)
        if refine and r["best_index"] >= 0:
            if opt.ascent_steps > 0:                 # projected-gradient variant (fixed number of steps)
                v = np.where(np.isnan(r["values"]), -np.inf, r["values"])
                top = np.argsort(-v, kind="stable")[:max(1, min(int(opt.ascent_top), v.size))]
                r2 = model.acquire_ascent(a.kind, a.params(), data[:, top], lb, ub, steps=int(opt.ascent_steps))
            else:
                r2 = _refine(opt, data, r["values"])
            if r2["best_index"] >= 0 and r2["best_value"] > r["best_value"]:
                r = dict(r, best_value=r2["best_value"], best_x=r2["best_x"])
    if opt.direct and str(opt.method).startswith("GN_DIRECT"):
        # the reference's derivative-free global search (:GN_DIRECT_L, maxeval evaluations; acquisition.jl:7-9,31-36) inside the library
        rd = model.acquire_direct(a.kind, a.params(), lb, ub, maxeval=int(opt.maxeval) if opt.maxeval else 2000, maxtime=float(opt.maxtime),
                                  width=int(opt.direct_width), seed=opt.seed + (1 << 32),
                                  variant=0 if "DIRECT_L" in str(opt.method) else 1)      # :GN_DIRECT_L* -> locally biased, :GN_DIRECT -> Jones' original
        if rd["best_index"] >= 0 and (r["best_index"] < 0 or rd["best_value"] > r["best_value"]):
            r = dict(r, best_value=rd["best_value"], best_x=rd["best_x"], best_index=rd["best_index"], best_from="direct")
    opt.seed += 1
    opt.last = r
    if r["best_index"] < 0:
        return -np.inf, lb                           # :55-56
    return r["best_value"], r["best_x"].copy()


def acquire_model_max(o, options=None):
    """acquire_model_max(o; options) (acquisition.jl:45-47): MaxMean on the current model."""
    options = o.acquisitionoptions if options is None else options
    return acquire_max(MaxMean(), o.model, o.lowerbounds, o.upperbounds, options)
