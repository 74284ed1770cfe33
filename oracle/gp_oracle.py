"""CPU restatement of the reference's GP-posterior + acquisition path.  TEST INFRASTRUCTURE ONLY.

*** PARITY UNPINNED at the GaussianProcesses.jl boundary ***
The reference (jbrea/BayesianOptimization.jl v0.2.5) is pure Julia and its arithmetic lives in
un-vendored packages that are absent from /root/reference and cannot be run in this image (no julia,
no network):  GaussianProcesses.jl (compat 0.9-0.12), ElasticPDMats.jl 0.2.3, NLopt.jl 0.4-1,
ForwardDiff 0.9/0.10, SpecialFunctions (Project.toml:20-31; no Manifest => no exact pin).
The reference's own tests hold no numeric golden vectors for mu / sigma^2 / acquisition values
(SURVEY.md section 4).  What *is* pinned here (tests/test_oracle.py):
  * test/acquisition.jl:11-12  -- 1-point GP, argmax of the posterior mean is x = 1.0
  * test/acquisitionfunctions.jl:8-10 -- batched call == per-point call, exactly
  * test/warmstart.jl:64 -- tau == maximum(y) after setparams!
  * analytic closed forms (1- and 2-point GPs) and mpmath 60-digit spot checks
This module is "a CPU restatement of the reference path", never "the reference".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this file.  The product path (bayesianoptimization.jl_b200) never does.

Conventions follow the reference: Float64 everywhere, points are COLUMNS (model.x is D x N,
src/models/gp.jl:9; candidates are D x M, test/acquisitionfunctions.jl:6).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.linalg import cholesky, solve_triangular
from scipy.special import erf

EPS = float(np.finfo(np.float64).eps)
LOG2PI = math.log(2.0 * math.pi)

KERNELS = ("SEIso", "SEArd", "Mat12Iso", "Mat12Ard", "Mat32Iso", "Mat32Ard", "Mat52Iso", "Mat52Ard")
MEANS = ("MeanZero", "MeanConst")
ACQS = ("PI", "EI", "UCB", "TS", "MI", "MaxMean")


# --------------------------------------------------------------------------------------------------
# kernels  (EXT GaussianProcesses.jl: SEArd/SEIso/Mat12/32/52; SURVEY.md Appendix A "Parametrisation")
# --------------------------------------------------------------------------------------------------
def kernel_is_iso(kind: str) -> bool:
    return kind.endswith("Iso")


def n_kernel_params(kind: str, D: int) -> int:
    return 2 if kernel_is_iso(kind) else D + 1


def _phi_psi(kind: str, r2: np.ndarray):
    """k = sf2 * phi(r2);  psi = -2 dphi/dr2, so  dk/dx*_d = -sf2 psi (x*_d - X_d)/l_d^2  and
    dk/dll_d = sf2 psi z_d^2  (z = Delta/l).  r2 = sum_d z_d^2."""
    fam = kind[:5]
    if fam.startswith("SE"):
        phi = np.exp(-0.5 * r2)
        return phi, phi
    r = np.sqrt(r2)
    if fam == "Mat12":
        e = np.exp(-r)
        with np.errstate(divide="ignore", invalid="ignore"):
            psi = np.where(r > 0, e / r, 0.0)
        return e, psi
    if fam == "Mat32":
        s = math.sqrt(3.0) * r
        e = np.exp(-s)
        return (1.0 + s) * e, 3.0 * e
    if fam == "Mat52":
        s = math.sqrt(5.0) * r
        e = np.exp(-s)
        return (1.0 + s + s * s / 3.0) * e, (5.0 / 3.0) * (1.0 + s) * e
    raise ValueError(kind)


def _scaled_sqdist(Xa: np.ndarray, Xb: np.ndarray, inv_ell: np.ndarray, chunk: int = 1024) -> np.ndarray:
    """r2[i,j] = sum_d ((Xa[d,i]-Xb[d,j]) / l_d)^2 by direct differences (no |a|^2+|b|^2-2ab cancellation)."""
    Za = (Xa * inv_ell[:, None]).T.copy()  # Na x D
    Zb = (Xb * inv_ell[:, None]).T.copy()  # Nb x D
    out = np.empty((Za.shape[0], Zb.shape[0]))
    for s in range(0, Za.shape[0], chunk):
        d = Za[s:s + chunk, None, :] - Zb[None, :, :]
        out[s:s + chunk] = np.einsum("ijd,ijd->ij", d, d)
    return out


@dataclass
class GPOracle:
    """GPE / ElasticGPE restated (EXT GaussianProcesses.jl; reference call sites src/models/gp.jl:2-18)."""
    D: int
    kernel: str = "SEArd"
    mean: str = "MeanConst"
    ll: np.ndarray = None          # log length-scales, length D (Ard) or 1 (Iso)
    lsigma: float = 0.0            # log signal std
    lognoise: float = -2.0
    beta: float = 0.0              # MeanConst parameter
    X: np.ndarray = field(default=None, repr=False)
    y: np.ndarray = field(default=None, repr=False)
    U: np.ndarray = field(default=None, repr=False)      # upper factor, Sigma = U^T U
    alpha: np.ndarray = field(default=None, repr=False)
    mll: float = float("nan")
    jitter_tries: int = 0

    def __post_init__(self):
        assert self.kernel in KERNELS and self.mean in MEANS
        nl = 1 if kernel_is_iso(self.kernel) else self.D
        self.ll = np.zeros(nl) if self.ll is None else np.atleast_1d(np.asarray(self.ll, float)).copy()
        assert self.ll.shape == (nl,)
        if self.X is None:
            self.X = np.zeros((self.D, 0))
            self.y = np.zeros(0)

    # -- parameters: theta = [logNoise; mean params; kernel params (ll..., lsigma)]  (gp.jl:55-58, App. A)
    def get_params(self, noise=True, domean=True, kern=True) -> np.ndarray:
        p = []
        if noise:
            p.append(self.lognoise)
        if domean and self.mean == "MeanConst":
            p.append(self.beta)
        if kern:
            p.extend(self.ll.tolist())
            p.append(self.lsigma)
        return np.array(p, float)

    def set_params(self, theta, noise=True, domean=True, kern=True):
        theta = np.asarray(theta, float)
        i = 0
        if noise:
            self.lognoise = float(theta[i]); i += 1
        if domean and self.mean == "MeanConst":
            self.beta = float(theta[i]); i += 1
        if kern:
            n = self.ll.size
            self.ll = theta[i:i + n].copy(); i += n
            self.lsigma = float(theta[i]); i += 1
        assert i == theta.size

    @property
    def inv_ell(self) -> np.ndarray:
        ie = np.exp(-self.ll)
        return np.full(self.D, ie[0]) if kernel_is_iso(self.kernel) else ie

    @property
    def sf2(self) -> float:
        return math.exp(2.0 * self.lsigma)

    def mean_at(self, n: int) -> np.ndarray:
        return np.full(n, self.beta if self.mean == "MeanConst" else 0.0)

    def cov(self, Xa: np.ndarray, Xb: np.ndarray) -> np.ndarray:
        phi, _ = _phi_psi(self.kernel, _scaled_sqdist(Xa, Xb, self.inv_ell))
        return self.sf2 * phi

    # -- fit  (App. A "Fit"; quirk 10: noise = exp(2 logNoise) + eps, make_posdef! jitter, upper factor)
    def fit(self, X: np.ndarray, y: np.ndarray):
        self.X = np.array(X, float, order="F").reshape(self.D, -1)
        self.y = np.array(y, float).ravel()
        N = self.y.size
        if N == 0:
            self.U = np.zeros((0, 0)); self.alpha = np.zeros(0); self.mll = 0.0
            return self
        S = self.cov(self.X, self.X)
        S[np.diag_indices(N)] += math.exp(2.0 * self.lognoise) + EPS
        self.jitter_tries = 0
        while True:
            try:
                self.U = cholesky(S, lower=False)
                break
            except np.linalg.LinAlgError:
                if self.jitter_tries >= 10:
                    raise
                S[np.diag_indices(N)] += 1e-6 * np.trace(S) / N
                self.jitter_tries += 1
        r = self.y - self.mean_at(N)
        z = solve_triangular(self.U, r, trans="T", lower=False)
        self.alpha = solve_triangular(self.U, z, lower=False)
        logdet = 2.0 * np.sum(np.log(np.diag(self.U)))
        self.mll = -0.5 * (r @ self.alpha + logdet + N * LOG2PI)
        return self

    def append(self, Xnew: np.ndarray, ynew: np.ndarray):
        """update!(model::GPE{<:ElasticArray}, x, y) = append!(model, x, y)  (gp.jl:11).  The rank-m
        extension U12 = U11^-T A12, U22 = chol(A22 - U12^T U12) equals a refit in exact arithmetic;
        the oracle refits (same result to rounding)."""
        Xn = np.asarray(Xnew, float).reshape(self.D, -1)
        return self.fit(np.hstack([self.X, Xn]), np.concatenate([self.y, np.ravel(ynew)]))

    # -- predict_f  (gp.jl:2-5,8; App. A "Predict"): latent variance, clamp max(.,0), column loop
    def predict(self, Xs: np.ndarray, return_aux: bool = False):
        Xs = np.asarray(Xs, float).reshape(self.D, -1)
        M = Xs.shape[1]
        N = self.y.size
        if N == 0:
            mu = self.mean_at(M); var = np.full(M, self.sf2)
            return (mu, var, None, None) if return_aux else (mu, var)
        Ks = self.cov(self.X, Xs)                                # N x M
        mu = self.mean_at(M) + Ks.T @ self.alpha
        V = solve_triangular(self.U, Ks, trans="T", lower=False)  # whiten!: v = U^-T k*
        var = np.maximum(self.sf2 - np.einsum("ij,ij->j", V, V), 0.0)
        return (mu, var, Ks, V) if return_aux else (mu, var)

    # -- rand(gp, X)  (reached from myrand(model, X::Matrix), gp.jl:7; App. A "Sample", quirk 9): EXT GaussianProcesses.jl rand!:
    #    mu, Sigma = predict_f(gp, X; full_cov = true); Sigma, chol = make_posdef!(Sigma); mu + unwhiten(PDMat(Sigma, chol), randn)
    #    i.e. mu + L eps with L the LOWER factor; make_posdef! retries add 1e-6 tr(Sigma)/M to the diagonal (<= 10 times).
    def posterior_cov(self, Xs: np.ndarray):
        Xs = np.asarray(Xs, float).reshape(self.D, -1)
        Kss = self.cov(Xs, Xs)
        if self.y.size == 0:
            return self.mean_at(Xs.shape[1]), Kss
        mu, _, _, V = self.predict(Xs, return_aux=True)
        return mu, Kss - V.T @ V

    def rand_joint(self, Xs: np.ndarray, eps: np.ndarray):
        mu, S = self.posterior_cov(Xs)
        M = mu.size
        tries = 0
        while True:
            try:
                Lc = cholesky(S, lower=True)
                break
            except np.linalg.LinAlgError:
                if tries >= 10:
                    raise
                S[np.diag_indices(M)] += 1e-6 * np.trace(S) / M
                tries += 1
        return mu + Lc @ np.asarray(eps, float), tries

    def predict_column_loop(self, Xs: np.ndarray):
        """Reference-shaped: one predict_full per column (GaussianProcesses.jl predict_f loop)."""
        Xs = np.asarray(Xs, float).reshape(self.D, -1)
        mu = np.empty(Xs.shape[1]); var = np.empty(Xs.shape[1])
        for k in range(Xs.shape[1]):
            ks = self.cov(self.X, Xs[:, k:k + 1])[:, 0]
            mu[k] = (self.beta if self.mean == "MeanConst" else 0.0) + ks @ self.alpha
            v = solve_triangular(self.U, ks, trans="T", lower=False)
            var[k] = max(self.sf2 - v @ v, 0.0)
        return mu, var

    # -- closed-form gradients of mu, sigma^2 wrt x* (replace ForwardDiff, acquisition.jl:13-15; App. A)
    def predict_grad(self, Xs: np.ndarray):
        Xs = np.asarray(Xs, float).reshape(self.D, -1)
        mu, var, Ks, V = self.predict(Xs, return_aux=True)
        W = solve_triangular(self.U, V, lower=False)             # w = Sigma^-1 k* = U^-1 v
        ie2 = self.inv_ell ** 2
        _, psi = _phi_psi(self.kernel, _scaled_sqdist(self.X, Xs, self.inv_ell))
        G = self.sf2 * psi                                        # N x M
        dmu = np.empty_like(Xs); dvar = np.empty_like(Xs)
        for d in range(self.D):
            dk = -G * (Xs[d][None, :] - self.X[d][:, None]) * ie2[d]     # dk*_i/dx*_d
            dmu[d] = np.einsum("ij,i->j", dk, self.alpha)
            dvar[d] = -2.0 * np.einsum("ij,ij->j", dk, W)
        dvar[:, var <= 0.0] = 0.0                                 # clamp active => zero gradient
        return mu, var, dmu, dvar

    # -- marginal likelihood + gradient (gp.jl:59-64 closure; App. A "dmll"), order [logNoise, mean, kernel]
    def mll_dmll(self, theta=None, noise=True, domean=True, kern=True):
        if theta is not None:
            self.set_params(theta, noise=noise, domean=domean, kern=kern)
        self.fit(self.X, self.y)
        N = self.y.size
        Uinv = solve_triangular(self.U, np.eye(N), lower=False)
        Sinv = Uinv @ Uinv.T
        A = np.outer(self.alpha, self.alpha) - Sinv
        g = []
        if noise:
            g.append(math.exp(2.0 * self.lognoise) * np.trace(A))
        if domean and self.mean == "MeanConst":
            g.append(float(np.sum(self.alpha)))
        if kern:
            ie = self.inv_ell
            r2 = _scaled_sqdist(self.X, self.X, ie)
            phi, psi = _phi_psi(self.kernel, r2)
            AG = A * (self.sf2 * psi)
            per_d = np.empty(self.D)
            for d in range(self.D):
                zd = (self.X[d][:, None] - self.X[d][None, :]) * ie[d]
                per_d[d] = 0.5 * np.sum(AG * zd * zd)
            if kernel_is_iso(self.kernel):
                g.append(float(per_d.sum()))
            else:
                g.extend(per_d.tolist())
            g.append(float(np.sum(A * (self.sf2 * phi))))       # d/dlsigma: dK = 2K -> 1/2 tr(A 2K)
        g = np.array(g, float)
        f = self.mll
        pri = getattr(self, "priors", None)          # EXT set_priors!: target = mll + sum of Normal log densities of the swept parameters
        if pri is not None:
            th = self.get_params(noise=noise, domean=domean, kern=kern)
            full = self.get_params()
            keep = []
            i = 0
            if noise: keep.append(i)
            i += 1
            if self.mean == "MeanConst":
                if domean: keep.append(i)
                i += 1
            if kern: keep.extend(range(i, full.size))
            for r, k in enumerate(keep):
                if pri[k] is not None:
                    z = (th[r] - pri[k][0]) / pri[k][1]
                    f += -0.5 * z * z - math.log(pri[k][1]) - 0.5 * math.log(2 * math.pi)
                    g[r] += -z / pri[k][1]
        return f, g


# --------------------------------------------------------------------------------------------------
# acquisition functors AS CODED (src/acquisitionfunctions.jl, src/utils.jl:48-49)
# --------------------------------------------------------------------------------------------------
def normal_pdf(mu, s2):          # utils.jl:48 -- N(0, s2) density, i.e. phi(z)/sigma
    return 1.0 / np.sqrt(2.0 * np.pi * s2) * np.exp(-mu ** 2 / (2.0 * s2))


def normal_cdf(mu, s2):          # utils.jl:49 -- via erf, not erfc
    return 0.5 * (1.0 + erf(mu / np.sqrt(2.0 * s2)))


def acq_value(kind: str, params, mu, s2, eps=None):
    """params: PI/EI -> (tau,), UCB -> (beta_t,), MI -> (sqrt_alpha, gamma_hat), MaxMean/TS -> ().
    TS needs eps ~ N(0,1) per candidate (independent per-candidate semantics, quirk 9)."""
    mu = np.asarray(mu, float); s2 = np.asarray(s2, float)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if kind == "PI":                                   # acquisitionfunctions.jl:24-27
            tau = params[0]
            return np.where(s2 == 0, (mu > tau).astype(float), normal_cdf(mu - tau, s2))
        if kind == "EI":                                   # :47-50  (Delta Phi(z) + phi(z), quirk 1)
            tau = params[0]
            v = (mu - tau) * normal_cdf(mu - tau, s2) + np.sqrt(s2) * normal_pdf(mu - tau, s2)
            return np.where(s2 == 0, np.where(mu > tau, mu - tau, 0.0), v)
        if kind == "UCB":                                  # :96
            return mu + params[0] * np.sqrt(s2)
        if kind == "MI":                                   # :141
            sa, gh = params
            return mu + sa * (np.sqrt(s2 + gh) - np.sqrt(gh))
        if kind == "MaxMean":                              # :111
            return mu.copy()
        if kind == "TS":                                   # :108 + gp.jl:6
            return mu + np.sqrt(s2) * np.asarray(eps, float)
    raise ValueError(kind)


def acq_partials(kind: str, params, mu, s2):
    """(da/dmu, da/ds2) of the functors AS CODED.  s2 == 0 -> both taken as the
    limit used by the kernels: da/dmu of the exact-zero branch, da/ds2 = 0."""
    mu = np.asarray(mu, float); s2 = np.asarray(s2, float)
    one = np.ones_like(mu); zero = np.zeros_like(mu)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        sig = np.sqrt(s2)
        if kind in ("PI", "EI"):
            tau = params[0]
            d = mu - tau
            z = d / sig
            ph = np.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)
            Ph = 0.5 * (1.0 + erf(z / math.sqrt(2.0)))
            if kind == "PI":
                return np.where(s2 == 0, zero, ph / sig), np.where(s2 == 0, zero, -z * ph / (2.0 * s2))
            amu = Ph + z * ph * (1.0 - 1.0 / sig)
            as2 = z * z * (1.0 - sig) * ph / (2.0 * s2)
            return np.where(s2 == 0, (mu > tau).astype(float), amu), np.where(s2 == 0, zero, as2)
        if kind == "UCB":
            return one, np.where(s2 == 0, zero, params[0] / (2.0 * sig))
        if kind == "MI":
            sa, gh = params
            den = np.sqrt(s2 + gh)
            return one, np.where(den == 0, zero, sa / (2.0 * den))
        if kind == "MaxMean":
            return one, zero
    raise ValueError(kind)


def acq_grad(gp: GPOracle, kind: str, params, Xs):
    """value and D x M gradient of a(mu(x), s2(x)) (closed form; replaces wrap_gradient, acquisition.jl:11-17)."""
    mu, var, dmu, dvar = gp.predict_grad(Xs)
    a = acq_value(kind, params, mu, var)
    amu, as2 = acq_partials(kind, params, mu, var)
    return a, amu[None, :] * dmu + as2[None, :] * dvar


def first_strict_argmax(values) -> int:
    """acquire_max's selection rule (acquisition.jl:55-66): running 'f > maxf' from -Inf keeps the FIRST
    strict maximum; NaN never wins; nothing wins -> -1 (the reference then returns lowerbounds)."""
    best, idx = -math.inf, -1
    for i, f in enumerate(np.asarray(values, float)):
        if f > best:
            best, idx = f, i
    return idx


def first_strict_argmax_np(values) -> int:
    v = np.asarray(values, float)
    ok = ~np.isnan(v) & (v > -np.inf)
    if not ok.any():
        return -1
    m = v[ok].max()
    return int(np.flatnonzero(ok & (v == m))[0])


# -- setparams! (acquisitionfunctions.jl:3,44-46,91-95,131-140) -------------------------------------
def maxy(gp: GPOracle) -> float:                         # gp.jl:10
    return -math.inf if gp.y.size == 0 else float(np.max(gp.y))


def brochu_beta(D: int, nobs: int, delta: float = 0.1) -> float:   # :88-95 (quirk 5)
    nobs = 1 if nobs == 0 else nobs
    return math.sqrt(2.0 * math.log(float(nobs) ** (D / 2.0 + 2.0) * math.pi ** 2 / (3.0 * delta)))


def mi_gamma_update(gp: GPOracle, gamma_hat: float) -> float:      # :131-140 (quirk 6)
    if gp.y.size == 0:
        return 0.0
    _, s2 = gp.predict(gp.X[:, -1:])
    return gamma_hat + float(s2[0])


# --------------------------------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller, keyed by (seed, global candidate index): sharding-invariant TS noise.
# NOT in the reference (it uses Julia's global RNG, gp.jl:6); this defines the repo's own stream and the
# CUDA kernel implements the identical generator so TS is checkable bit-for-bit up to libm ulps.
# --------------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: (n,4) uint32, key: (n,2) uint32 -> (n,4) uint32."""
    c = ctr.astype(np.uint32).copy(); k = key.astype(np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = _M0 * c[:, 0].astype(np.uint64)
        p1 = _M1 * c[:, 2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
        c = np.stack([hi1 ^ c[:, 1] ^ k[:, 0], lo1, hi0 ^ c[:, 3] ^ k[:, 1], lo0], axis=1)
        with np.errstate(over="ignore"):
            k = np.stack([k[:, 0] + _W0, k[:, 1] + _W1], axis=1)
    return c


def philox_normal(seed: int, idx: np.ndarray) -> np.ndarray:
    """eps_i ~ N(0,1) for global candidate indices idx (int64).  counter = (idx_lo, idx_hi, 0, 0),
    key = (seed_lo, seed_hi); u1 = ((r0:r1 >> 11) + 1) * 2^-53 in (0,1], u2 = (r2:r3 >> 11) * 2^-53 in
    [0,1); eps = sqrt(-2 ln u1) cos(2 pi u2)."""
    idx = np.asarray(idx, np.uint64)
    n = idx.size
    ctr = np.zeros((n, 4), np.uint32)
    ctr[:, 0] = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (idx >> np.uint64(32)).astype(np.uint32)
    key = np.zeros((n, 2), np.uint32)
    s = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
    key[:, 0] = np.uint32(s & np.uint64(0xFFFFFFFF)); key[:, 1] = np.uint32(s >> np.uint64(32))
    r = philox4x32_10(ctr, key).astype(np.uint64)
    a = (r[:, 0] << np.uint64(32)) | r[:, 1]
    b = (r[:, 2] << np.uint64(32)) | r[:, 3]
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * 2.0 ** -53
    u2 = (b >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


# --------------------------------------------------------------------------------------------------
# candidate generator (src/utils.jl:101-120 latin_hypercube_sampling): per dim n jittered strata, shuffled.
# --------------------------------------------------------------------------------------------------
def latin_hypercube_sampling(mins, maxs, n: int, rng: np.random.Generator) -> np.ndarray:
    mins = np.asarray(mins, float); maxs = np.asarray(maxs, float)
    if mins.size != maxs.size:
        raise ValueError("mins and maxs should have the same length")
    if not np.all(mins <= maxs):
        raise ValueError("mins[i] should not exceed maxs[i]")
    out = np.zeros((mins.size, n))
    for i in range(mins.size):
        step = (maxs[i] - mins[i]) / n
        cube = mins[i] + step * (np.arange(n) + rng.random(n))
        rng.shuffle(cube)
        out[i, :] = cube
    return out


# --------------------------------------------------------------------------------------------------
# device-side candidate generation / batched ascent restated (csrc/search.cu).  These are the REPO's own algorithms for
# the SURVEY 8f-2/8f-3 rows (the reference uses Random.shuffle! and NLopt); the oracle restates them to check the kernels.
# --------------------------------------------------------------------------------------------------
def _philox_words(c, k):
    return philox4x32_10(np.asarray(c, np.uint32).reshape(-1, 4), np.asarray(k, np.uint32).reshape(-1, 2))


def _perm_index(x: np.ndarray, n: int, bits: int, key) -> np.ndarray:
    mask = np.uint64((1 << bits) - 1)
    sh = np.uint64(bits // 2 + 1)
    k = [np.uint64(int(v)) for v in key]
    x = x.astype(np.uint64).copy()
    todo = np.ones(x.size, bool)
    with np.errstate(over="ignore"):
        while todo.any():
            y = x[todo]
            y = (y + k[0]) & mask
            y = (y * (np.uint64(2) * k[1] + np.uint64(1))) & mask
            y ^= y >> sh
            y = (y * (np.uint64(2) * k[2] + np.uint64(1))) & mask
            y ^= y >> sh
            y = (y + k[3]) & mask
            x[todo] = y
            todo[todo] = y >= np.uint64(n)
    return x


def lhs_device(lb, ub, n_total: int, seed: int, offset: int = 0, n_local: int = None) -> np.ndarray:
    """csrc/search.cu:lhs_kernel restated: stratum = keyed permutation of the global column index, Philox jitter."""
    lb = np.asarray(lb, float); ub = np.asarray(ub, float)
    D = lb.size
    n_local = n_total - offset if n_local is None else n_local
    bits = 1
    while bits < 63 and (1 << bits) < n_total:
        bits += 1
    j = np.arange(offset, offset + n_local, dtype=np.uint64)
    s_lo, s_hi = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    out = np.empty((D, n_local))
    for d in range(D):
        key = _philox_words([[d, 0x4C485321, 0, 0]], [[s_lo, s_hi]])[0]
        stratum = _perm_index(j, n_total, bits, key)
        ctr = np.stack([(j & np.uint64(0xFFFFFFFF)).astype(np.uint32), (j >> np.uint64(32)).astype(np.uint32),
                        np.full(n_local, d, np.uint32), np.full(n_local, 0x4A495454, np.uint32)], axis=1)
        u = philox4x32_10(ctr, np.tile(np.array([[s_lo, s_hi]], np.uint32), (n_local, 1))).astype(np.uint64)
        jit = (((u[:, 0] << np.uint64(32)) | u[:, 1]) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        step = (ub[d] - lb[d]) / float(n_total)
        out[d] = lb[d] + step * (stratum.astype(np.float64) + jit)
    return out


def ascent(gp: "GPOracle", kind: str, params, X0, lb, ub, steps: int = 20, step0: float = 0.05):
    """csrc/search.cu:ascent_step_kernel restated with the oracle's value/gradient."""
    lb = np.asarray(lb, float); ub = np.asarray(ub, float); rng = ub - lb
    X = np.array(X0, float).reshape(lb.size, -1).copy()
    M = X.shape[1]
    Xb = X.copy(); Fb = np.full(M, -np.inf); Gb = np.zeros_like(X); S = np.full(M, step0)
    for it in range(steps + 1):
        v, g = acq_grad(gp, kind, params, X)
        better = np.ones(M, bool) if it == 0 else (v > Fb)
        Fb = np.where(better, v, Fb)
        Xb[:, better] = X[:, better]; Gb[:, better] = g[:, better]
        if it > 0:
            S = np.where(better, np.minimum(S * 1.3, 0.5), S * 0.35)
        if it == steps:
            return Xb, Fb
        gu = Gb * rng[:, None]
        nrm = np.sqrt(np.sum(gu * gu, axis=0))
        with np.errstate(divide="ignore", invalid="ignore"):
            stepv = np.where((nrm > 0) & np.isfinite(nrm), rng[:, None] * S * (gu / nrm), 0.0)
        X = np.clip(Xb + stepv, lb[:, None], ub[:, None])


# --------------------------------------------------------------------------------------------------
# test functions used by the BASELINE configs (test/branin.jl:1-5, examples/branin_hartmann.jl:12-22)
# --------------------------------------------------------------------------------------------------
def branin(x1, x2, a=1.0, b=5.1 / (4 * math.pi ** 2), c=5 / math.pi, r=6.0, s=10.0, t=1 / (8 * math.pi)):
    return a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s


_H_ALPHA = np.array([1.0, 1.2, 3.0, 3.2])
_H_A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]], float)
_H_P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                        [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]], float)


def hartmann6(X: np.ndarray) -> np.ndarray:
    X = np.asarray(X, float).reshape(6, -1)
    d = X.T[:, None, :] - _H_P[None, :, :]
    return -np.sum(_H_ALPHA[None, :] * np.exp(-np.sum(_H_A[None, :, :] * d * d, axis=2)), axis=1)
