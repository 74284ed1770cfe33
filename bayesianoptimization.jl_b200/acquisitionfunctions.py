"""Acquisition types of the reference (src/acquisitionfunctions.jl), kept as thin host objects.

The functor types, their fields, `setparams!` and `acquisitionfunction` keep the reference's names and meaning.
The scalar formulae below are host glue (used by `setparams!` and for single evaluations); over a candidate
matrix the same formulae run fused on the B200 inside b200bo_acquire (csrc/acq.cu: acq_eval).
"""
from __future__ import annotations

import math

import numpy as np

from .gp import B200GPE, dims, maxy, mean_var


class AbstractAcquisition:
    kind = None                       # C-ABI acquisition kind

    def params(self):
        return ()

    def setparams(self, model):       # setparams!(a, model) = nothing   (:3)
        return None


def normal_pdf(mu, s2):               # src/utils.jl:48 (N(0, s2) density -- quirk 1)
    return 1.0 / math.sqrt(2.0 * math.pi * s2) * math.exp(-mu ** 2 / (2.0 * s2))


def normal_cdf(mu, s2):               # src/utils.jl:49
    return 0.5 * (1.0 + math.erf(mu / math.sqrt(2.0 * s2)))


class ProbabilityOfImprovement(AbstractAcquisition):      # :21-28
    kind = "PI"

    def __init__(self, tau: float = -math.inf):
        self.tau = float(tau)

    def params(self):
        return (self.tau,)

    def setparams(self, model):                            # :44-46 (monotone tau, quirk 4)
        self.tau = max(maxy(model), self.tau)

    def __call__(self, mu, s2):
        if s2 == 0:
            return float(mu > self.tau)
        return normal_cdf(mu - self.tau, s2)


class ExpectedImprovement(AbstractAcquisition):            # :40-50
    kind = "EI"

    def __init__(self, tau: float = -math.inf):
        self.tau = float(tau)

    def params(self):
        return (self.tau,)

    def setparams(self, model):
        self.tau = max(maxy(model), self.tau)

    def __call__(self, mu, s2):
        if s2 == 0:
            return mu - self.tau if mu > self.tau else 0.0
        return (mu - self.tau) * normal_cdf(mu - self.tau, s2) + math.sqrt(s2) * normal_pdf(mu - self.tau, s2)


class BetaScaling:
    pass


class BrochuBetaScaling(BetaScaling):                      # :66-68
    def __init__(self, delta: float = 0.1):
        self.delta = float(delta)


class NoBetaScaling(BetaScaling):                          # :72
    pass


class UpperConfidenceBound(AbstractAcquisition):           # :81-96
    kind = "UCB"

    def __init__(self, scaling: BetaScaling = None, beta_t: float = 1.0):
        self.scaling = BrochuBetaScaling(0.1) if scaling is None else scaling
        self.beta_t = float(beta_t)

    def params(self):
        return (self.beta_t,)

    def setparams(self, model):                            # :91-95 (quirk 5)
        if isinstance(self.scaling, BrochuBetaScaling):
            D, nobs = dims(model)
            nobs = 1 if nobs == 0 else nobs
            self.beta_t = math.sqrt(2.0 * math.log(float(nobs) ** (D / 2.0 + 2.0) * math.pi ** 2 / (3.0 * self.scaling.delta)))

    def __call__(self, mu, s2):
        return mu + self.beta_t * math.sqrt(s2)


class ThompsonSamplingSimple(AbstractAcquisition):         # :107-108
    kind = "TS"


class MaxMean(AbstractAcquisition):                        # :110-111
    kind = "MaxMean"

    def __call__(self, mu, s2):
        return mu


class MutualInformation(AbstractAcquisition):              # :126-141
    kind = "MI"

    def __init__(self, alpha: float = 1.0, gamma_hat: float = 0.0):
        self.sqrt_alpha = math.sqrt(alpha)
        self.gamma_hat = float(gamma_hat)

    def params(self):
        return (self.sqrt_alpha, self.gamma_hat)

    def setparams(self, model):                            # :131-140 (stateful, quirk 6)
        D, nobs = dims(model)
        if nobs == 0:
            self.gamma_hat = 0.0
        else:
            _, s2 = mean_var(model, model.x[:, -1])
            self.gamma_hat += s2

    def __call__(self, mu, s2):
        return mu + self.sqrt_alpha * (math.sqrt(s2 + self.gamma_hat) - math.sqrt(self.gamma_hat))


def setparams(a: AbstractAcquisition, model):
    """setparams!(a, model)."""
    return a.setparams(model)


def acquisitionfunction(a: AbstractAcquisition, model: B200GPE, seed: int = 0):
    """acquisitionfunction(a, model) (:4-9, :108, :111): closure over a Vector (-> scalar) or a D x M Matrix
    (-> length-M vector); the matrix form is one fused launch."""
    def f(x):
        x = np.asarray(x, float)
        r = model.acquire(a.kind, a.params(), x, seed=seed)
        return float(r["values"][0]) if x.ndim == 1 else r["values"]
    return f
