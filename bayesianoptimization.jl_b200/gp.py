"""Model adapter: the reference's model plugin API (src/models/gp.jl) on top of libb200bo.

`B200GPE` stands where the reference uses `ElasticGPE(D; mean, kernel, logNoise, capacity)` (README.md:22-26):
it implements the generic functions the BO loop dispatches on -- mean_var, myrand, dims, maxy, update!,
optimizemodel! -- by calling the C ABI (same calls as julia/B200BayesOpt.jl makes with ccall).  Points are
COLUMNS (D x N) as in the reference (gp.jl:9).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import lib, check, dptr


# --- EXT GaussianProcesses.jl constructors used in reference scripts (README.md:22-26, test/branin.jl:24-26) ---
class MeanZero:
    kind = "MeanZero"
    params = ()


class MeanConst:
    kind = "MeanConst"

    def __init__(self, beta: float = 0.0):
        self.beta = float(beta)

    @property
    def params(self):
        return (self.beta,)


class _Kernel:
    def __init__(self, kind, ll, lsigma):
        self.kind = kind
        self.ll = np.atleast_1d(np.asarray(ll, float)).copy()
        self.lsigma = float(lsigma)


def SEIso(ll, lsigma): return _Kernel("SEIso", [ll], lsigma)
def SEArd(ll, lsigma): return _Kernel("SEArd", ll, lsigma)
def Mat12Iso(ll, lsigma): return _Kernel("Mat12Iso", [ll], lsigma)
def Mat12Ard(ll, lsigma): return _Kernel("Mat12Ard", ll, lsigma)
def Mat32Iso(ll, lsigma): return _Kernel("Mat32Iso", [ll], lsigma)
def Mat32Ard(ll, lsigma): return _Kernel("Mat32Ard", ll, lsigma)
def Mat52Iso(ll, lsigma): return _Kernel("Mat52Iso", [ll], lsigma)
def Mat52Ard(ll, lsigma): return _Kernel("Mat52Ard", ll, lsigma)


class B200GPE:
    """ElasticGPE-shaped GP whose factor, alpha and data live in B200 HBM."""

    def __init__(self, D: int, mean=None, kernel=None, logNoise: float = -2.0, capacity: int = 3000, device: int = 0, n_gpus=None,
                 devices=None):
        mean = MeanZero() if mean is None else mean
        kernel = SEArd(np.zeros(D), 0.0) if kernel is None else kernel
        if not kernel.kind.endswith("Iso") and kernel.ll.size != D:
            raise ValueError("ARD kernel needs one length-scale per input dimension")
        self.D = int(D)
        self.mean_kind, self.kernel_kind = mean.kind, kernel.kind
        self._h = C.c_void_p()
        if n_gpus is None and devices is None:
            check(lib.b200bo_create(C.byref(self._h), device, self.D, int(capacity), _lib.KERNEL_KINDS[kernel.kind],
                                    _lib.MEAN_KINDS[mean.kind]))
        else:       # ONE process, several GPUs: replicas of the model behind one handle (b200bo_create_multi)
            devs = list(devices) if devices is not None else list(range(int(n_gpus)))
            arr = (C.c_int32 * len(devs))(*devs)
            check(lib.b200bo_create_multi(C.byref(self._h), len(devs), arr, self.D, int(capacity), _lib.KERNEL_KINDS[kernel.kind],
                                          _lib.MEAN_KINDS[mean.kind]))
        theta = [float(logNoise), *mean.params, *kernel.ll.tolist(), kernel.lsigma]
        self.set_params(np.array(theta, float))

    @classmethod
    def from_data(cls, X, y, mean=None, kernel=None, logNoise: float = -2.0, **kw):
        """GPE(x, y, mean, kernel, logNoise) (test/acquisitionfunctions.jl:4, test/acquisition.jl:2)."""
        X = np.asarray(X, float)
        X = X.reshape(1, -1) if X.ndim == 1 else X
        gp = cls(X.shape[0], mean=mean, kernel=kernel, logNoise=logNoise, capacity=max(X.shape[1], 1), **kw)
        gp.fit(X, y)
        return gp

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.b200bo_destroy(h)

    # -- hyper-parameters: theta = [logNoise, (beta), ll..., lsigma] ------------------------------------------
    @property
    def num_params(self) -> int:
        p = C.c_int32()
        check(lib.b200bo_num_params(self._h, C.byref(p)), self._h)
        return p.value

    def get_params(self) -> np.ndarray:
        th = np.empty(self.num_params)
        check(lib.b200bo_get_params(self._h, dptr(th), th.size), self._h)
        return th

    def set_params(self, theta):
        th = np.ascontiguousarray(theta, float)
        check(lib.b200bo_set_params(self._h, dptr(th), th.size), self._h)

    def set_priors(self, priors):
        """EXT set_priors! (reference src/models/gp.jl:30-35): one entry per parameter in the order of get_params(): None = flat,
        (mu, sigma) = Normal.  The MAP objective (mll_sweep / map_fit) becomes mll + log prior.  set_priors(None) clears."""
        if priors is None:
            check(lib.b200bo_set_priors(self._h, 0, None, None, None), self._h)
            return
        P = self.num_params
        if len(priors) != P:
            raise ValueError("one prior per parameter")
        kind = (C.c_int32 * P)(*[0 if p is None else 1 for p in priors])
        a = np.array([0.0 if p is None else p[0] for p in priors], float)
        b = np.array([1.0 if p is None else p[1] for p in priors], float)
        check(lib.b200bo_set_priors(self._h, P, kind, dptr(a), dptr(b)), self._h)

    # -- data ------------------------------------------------------------------------------------------------
    def fit(self, X, y):
        X = np.asfortranarray(np.asarray(X, float).reshape(self.D, -1))
        y = np.ascontiguousarray(y, float).ravel()
        if X.shape[1] != y.size:
            raise ValueError("X must be D x N with N == length(y)")
        check(lib.b200bo_fit(self._h, dptr(X), dptr(y), y.size), self._h)

    def append(self, X, y):
        X = np.asfortranarray(np.asarray(X, float).reshape(self.D, -1))
        y = np.ascontiguousarray(y, float).ravel()
        if X.shape[1] != y.size:
            raise ValueError("X must be D x m with m == length(y)")
        check(lib.b200bo_append(self._h, dptr(X), dptr(y), y.size), self._h)

    @property
    def nobs(self) -> int:
        n = C.c_int64()
        check(lib.b200bo_dims(self._h, None, C.byref(n)), self._h)
        return n.value

    @property
    def x(self) -> np.ndarray:          # model.x, D x N
        X = np.empty((self.D, self.nobs), order="F")
        check(lib.b200bo_get_data(self._h, dptr(X), None), self._h)
        return X

    @property
    def y(self) -> np.ndarray:          # model.y
        y = np.empty(self.nobs)
        check(lib.b200bo_get_data(self._h, None, dptr(y)), self._h)
        return y

    @property
    def mll(self) -> float:
        v = C.c_double()
        check(lib.b200bo_get_mll(self._h, C.byref(v)), self._h)
        return v.value

    @property
    def alpha(self) -> np.ndarray:
        a = np.empty(self.nobs)
        check(lib.b200bo_get_alpha(self._h, dptr(a)), self._h)
        return a

    @property
    def factor(self) -> np.ndarray:     # upper U, Sigma = U'U
        n = self.nobs
        U = np.empty((n, n), order="F")
        check(lib.b200bo_get_factor(self._h, dptr(U)), self._h)
        return U

    @property
    def jitter_tries(self) -> int:
        t = C.c_int32()
        check(lib.b200bo_jitter_tries(self._h, C.byref(t)), self._h)
        return t.value

    def kmat(self) -> np.ndarray:
        n = self.nobs
        K = np.empty((n, n), order="F")
        check(lib.b200bo_kmat(self._h, dptr(K)), self._h)
        return K

    def timing_ms(self, which: int) -> float:
        ms = C.c_float()
        check(lib.b200bo_last_timing_ms(self._h, which, C.byref(ms)), self._h)
        return ms.value

    def fp64_peak_tflops(self) -> float:
        v = C.c_double()
        check(lib.b200bo_fp64_peak_tflops(self._h, C.byref(v)), self._h)
        return v.value

    def set_syrk_engine(self, engine: int) -> None:
        """1 = tcgen05 int8-slice trailing updates (default), 0 = DMMA tile GEMM; takes effect at the next fit."""
        check(lib.b200bo_set_syrk_engine(self._h, int(engine)), self._h)

    def i8_peak_tops(self) -> float:
        v = C.c_double()
        check(lib.b200bo_i8_peak_tops(self._h, C.byref(v)), self._h)
        return v.value

    def set_knob(self, name: str, value: int) -> None:
        check(lib.b200bo_set_knob(self._h, name.encode(), int(value)), self._h)

    @property
    def num_gpus(self) -> int:
        n = C.c_int32()
        check(lib.b200bo_num_gpus(self._h, C.byref(n)), self._h)
        return n.value

    def comm_init_rank(self, world: int, rank: int, unique_id: bytes) -> None:
        """one process per GPU: attach an NCCL communicator over the `world` handles that hold replicas of this model; from then on
        acquire() / b200bo_acquire_dev return the GLOBAL best (collective call)."""
        assert len(unique_id) == 128
        check(lib.b200bo_comm_init_rank(self._h, int(world), int(rank), unique_id), self._h)

    def comm_destroy(self) -> None:
        check(lib.b200bo_comm_destroy(self._h), self._h)

    def set_acq_engine(self, engine: int) -> None:
        """1 = acquisition step as an int8-slice GEMM against L^-1 on tcgen05 (default), 0 = blocked DMMA solves (A/B tests)."""
        check(lib.b200bo_set_acq_engine(self._h, int(engine)), self._h)

    @property
    def launch_count(self) -> int:
        n = C.c_int64()
        check(lib.b200bo_launch_count(self._h, C.byref(n)), self._h)
        return n.value

    # -- posterior / acquisition ------------------------------------------------------------------------------
    def _cands(self, X):
        X = np.asarray(X, float)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        if X.shape[0] != self.D:
            raise ValueError(f"candidates must be {self.D} x M")
        return np.asfortranarray(X)

    def predict(self, X):
        Xs = self._cands(X)
        M = Xs.shape[1]
        mu, var = np.empty(M), np.empty(M)
        check(lib.b200bo_predict(self._h, dptr(Xs), M, dptr(mu), dptr(var)), self._h)
        return mu, var

    def rand_joint(self, X, seed: int = 0, idx_offset: int = 0):
        """ONE joint posterior draw over the columns of X (EXT rand(gp, X), reached from myrand(model, X::Matrix), gp.jl:7).
        Returns dict(sample, mu, tries)."""
        Xs = self._cands(X)
        M = Xs.shape[1]
        out, mu = np.empty(M), np.empty(M)
        tries = C.c_int32()
        check(lib.b200bo_rand_joint(self._h, dptr(Xs), M, seed & 0xFFFFFFFFFFFFFFFF, idx_offset, dptr(out), dptr(mu), C.byref(tries)), self._h)
        return dict(sample=out, mu=mu, tries=tries.value)

    def acquire(self, kind: str, params, X, seed: int = 0, idx_offset: int = 0, want_values=True, want_grad=False,
                want_mu_var=False, values_out=None, grad_out=None):
        """One fused acquisition step over the columns of X.  Returns a dict with best_value, best_index, best_x and
        the optional per-candidate arrays.  values_out (M) / grad_out (D x M, F-order) let the caller supply the output buffers: with
        PINNED buffers (candidates included) the library overlaps the PCIe transfers with the kernels chunk by chunk."""
        Xs = self._cands(X)
        M = Xs.shape[1]
        p = np.ascontiguousarray(params, float).ravel()
        vals = (values_out if values_out is not None else np.empty(M)) if want_values else None
        grad = (grad_out if grad_out is not None else np.empty((self.D, M), order="F")) if want_grad else None
        if vals is not None and vals.shape != (M,):
            raise ValueError("values_out must have M elements")
        if grad is not None and (grad.shape != (self.D, M) or not grad.flags.f_contiguous):
            raise ValueError("grad_out must be D x M in column-major order")
        mu = np.empty(M) if want_mu_var else None
        var = np.empty(M) if want_mu_var else None
        best = _lib.Best()
        bx = np.full(self.D, np.nan)
        check(lib.b200bo_acquire(self._h, _lib.ACQ_KINDS[kind], dptr(p) if p.size else None, p.size, dptr(Xs), M,
                                 seed & 0xFFFFFFFFFFFFFFFF, idx_offset, dptr(vals), dptr(grad), dptr(mu), dptr(var),
                                 C.byref(best), dptr(bx)), self._h)
        return dict(best_value=best.value, best_index=best.index, best_x=bx, values=vals, grad=grad, mu=mu, var=var)

    def lhs(self, lb, ub, n_total: int, seed: int = 0, offset: int = 0, n_local: int = None) -> np.ndarray:
        """latin_hypercube_sampling(lb, ub, n) (src/utils.jl:101-120) generated on the device: columns
        [offset, offset + n_local) of one global n_total-point design."""
        n_local = n_total - offset if n_local is None else n_local
        lb = np.ascontiguousarray(lb, float); ub = np.ascontiguousarray(ub, float)
        if lb.size != self.D or ub.size != self.D:
            raise ValueError("mins and maxs should have the same length")
        X = np.empty((self.D, n_local), order="F")
        check(lib.b200bo_lhs(self._h, dptr(lb), dptr(ub), n_total, offset, n_local, seed & 0xFFFFFFFFFFFFFFFF, dptr(X)), self._h)
        return X

    def acquire_lhs(self, kind: str, params, lb, ub, n_total: int, lhs_seed: int = 0, ts_seed: int = 0, offset: int = 0,
                    n_local: int = None, want_values=False):
        """LHS candidates born in HBM + one fused sweep (ScaledLHSIterator + acquire_max's loop, acquisition.jl:57-66)."""
        n_local = n_total - offset if n_local is None else n_local
        lb = np.ascontiguousarray(lb, float); ub = np.ascontiguousarray(ub, float)
        p = np.ascontiguousarray(params, float).ravel()
        vals = np.empty(n_local) if want_values else None
        best = _lib.Best(); bx = np.full(self.D, np.nan)
        check(lib.b200bo_acquire_lhs(self._h, _lib.ACQ_KINDS[kind], dptr(p) if p.size else None, p.size, dptr(lb), dptr(ub), n_total,
                                     offset, n_local, lhs_seed & 0xFFFFFFFFFFFFFFFF, ts_seed & 0xFFFFFFFFFFFFFFFF, dptr(vals),
                                     C.byref(best), dptr(bx)), self._h)
        return dict(best_value=best.value, best_index=best.index, best_x=bx, values=vals)

    def sobol(self, lb, ub, index0: int, n: int) -> np.ndarray:
        """points index0 .. index0+n-1 of the unscrambled Joe-Kuo Sobol sequence, generated on the device and scaled to [lb, ub]
        (ScaledSobolIterator, src/utils.jl:64-87).  D x n."""
        lb = np.ascontiguousarray(lb, float); ub = np.ascontiguousarray(ub, float)
        out = np.empty((self.D, int(n)), order="F")
        check(lib.b200bo_sobol(self._h, dptr(lb), dptr(ub), int(index0), int(n), dptr(out)), self._h)
        return out

    def acquire_lbfgs(self, kind: str, params, X, lb, ub, maxeval: int = 2000, ftol_rel: float = 0.0, ftol_abs: float = 0.0,
                      xtol_rel: float = 0.0, xtol_abs: float = 0.0, maxtime: float = 0.0, step0: float = 0.02, idx_offset: int = 0):
        """box-bounded L-BFGS ascents from ALL columns of X in lock-step on the fused value+gradient launch -- what the reference's
        per-restart NLopt :LD_LBFGS run does (src/acquisition.jl:59), with its stopping options (:24-27)."""
        Xs = self._cands(X)
        M = Xs.shape[1]
        lb = np.ascontiguousarray(lb, float); ub = np.ascontiguousarray(ub, float)
        p = np.ascontiguousarray(params, float).ravel()
        Xout = np.empty((self.D, M), order="F"); vals = np.empty(M); evals = np.empty(M)
        best = _lib.Best(); bx = np.full(self.D, np.nan)
        check(lib.b200bo_acquire_lbfgs(self._h, _lib.ACQ_KINDS[kind], dptr(p) if p.size else None, p.size, dptr(Xs), M, dptr(lb), dptr(ub),
                                       int(maxeval), float(ftol_rel), float(ftol_abs), float(xtol_rel), float(xtol_abs), float(maxtime),
                                       float(step0), idx_offset, dptr(Xout), dptr(vals), dptr(evals), C.byref(best), dptr(bx)), self._h)
        return dict(X=Xout, values=vals, evals=evals.astype(int), best_value=best.value, best_index=best.index, best_x=bx)

    def acquire_direct(self, kind: str, params, lb, ub, maxeval: int = 2000, maxtime: float = 0.0, width: int = 1, seed: int = 0,
                       want_trace: bool = False, variant: int = 0):
        """NLopt :GN_DIRECT_L (the reference's default search for ThompsonSamplingSimple, src/acquisition.jl:7-9) as a batched
        locally-biased DIRECT inside the library: one fused launch per iteration over all new rectangle centres.  variant 1 = Jones'
        original DIRECT (NLopt :GN_DIRECT)."""
        lb = np.ascontiguousarray(lb, float); ub = np.ascontiguousarray(ub, float)
        if lb.size != self.D or ub.size != self.D:
            raise ValueError("bounds must have length D")
        p = np.ascontiguousarray(params, float).ravel()
        Xt = np.full((self.D, int(maxeval)), np.nan, order="F") if want_trace else None
        ft = np.full(int(maxeval), np.nan) if want_trace else None
        ev = C.c_int32(); nb = C.c_int32(); best = _lib.Best(); bx = np.full(self.D, np.nan)
        check(lib.b200bo_acquire_direct(self._h, _lib.ACQ_KINDS[kind], dptr(p) if p.size else None, p.size, dptr(lb), dptr(ub), int(maxeval),
                                        float(maxtime), int(width), int(variant), seed & 0xFFFFFFFFFFFFFFFF, dptr(Xt), dptr(ft), C.byref(ev), C.byref(nb),
                                        C.byref(best), dptr(bx)), self._h)
        r = dict(best_value=best.value, best_index=best.index, best_x=bx, evals=ev.value, batches=nb.value)
        if want_trace:
            r.update(X=Xt[:, :ev.value], values=ft[:ev.value])
        return r

    def map_fit(self, theta0, lb, ub, noise=True, domean=True, kern=True, maxeval: int = 500, ftol_rel: float = 0.0, ftol_abs: float = 0.0,
                xtol_rel: float = 0.0, xtol_abs: float = 0.0, maxtime: float = 0.0):
        """optimizemodel!(::MAPGPOptimizer, model) (src/models/gp.jl:54-77) inside the library: L-BFGS ascent of mll over the selected
        parameters within [lb, ub] from the columns of theta0 (P x R starts in lock-step).  Leaves the model at the optimum."""
        T0 = np.asfortranarray(np.asarray(theta0, float))
        T0 = T0.reshape(-1, 1) if T0.ndim == 1 else T0
        P, R = T0.shape
        lb = np.ascontiguousarray(lb, float); ub = np.ascontiguousarray(ub, float)
        mask = (_lib.MASK_NOISE if noise else 0) | (_lib.MASK_MEAN if domean else 0) | (_lib.MASK_KERN if kern else 0)
        th = np.empty(P); mll = C.c_double(); ev = C.c_int32(); st = C.c_int32()
        check(lib.b200bo_map_fit(self._h, dptr(T0), P, R, mask, dptr(lb), dptr(ub), int(maxeval), float(ftol_rel), float(ftol_abs), float(xtol_rel),
                                 float(xtol_abs), float(maxtime), dptr(th), C.byref(mll), C.byref(ev), C.byref(st)), self._h)
        return dict(theta=th, mll=mll.value, evals=ev.value, status=st.value)

    def acquire_ascent(self, kind: str, params, X, lb, ub, steps: int = 20, step0: float = 0.05, idx_offset: int = 0):
        """M box-constrained gradient ascents in lock-step on the fused value+gradient kernel (acquisition.jl:59)."""
        Xs = self._cands(X)
        M = Xs.shape[1]
        lb = np.ascontiguousarray(lb, float); ub = np.ascontiguousarray(ub, float)
        p = np.ascontiguousarray(params, float).ravel()
        Xout = np.empty((self.D, M), order="F"); vals = np.empty(M)
        best = _lib.Best(); bx = np.full(self.D, np.nan)
        check(lib.b200bo_acquire_ascent(self._h, _lib.ACQ_KINDS[kind], dptr(p) if p.size else None, p.size, dptr(Xs), M, dptr(lb),
                                        dptr(ub), steps, step0, idx_offset, dptr(Xout), dptr(vals), C.byref(best), dptr(bx)), self._h)
        return dict(best_value=best.value, best_index=best.index, best_x=bx, values=vals, X=Xout)

    def mll_sweep(self, Theta, noise=True, domean=True, kern=True, want_grad=True):
        Theta = np.asfortranarray(np.asarray(Theta, float))
        Theta = Theta.reshape(-1, 1) if Theta.ndim == 1 else Theta
        P, S = Theta.shape
        mask = (_lib.MASK_NOISE if noise else 0) | (_lib.MASK_MEAN if domean else 0) | (_lib.MASK_KERN if kern else 0)
        mll = np.empty(S)
        dmll = np.empty((P, S), order="F") if want_grad else None
        check(lib.b200bo_mll_sweep(self._h, dptr(Theta), P, S, mask, dptr(mll), dptr(dmll)), self._h)
        return mll, dmll


# ------------------------------------------------------------------------------------------------------------
# the generic functions of src/models/gp.jl, same names and argument meaning
# ------------------------------------------------------------------------------------------------------------
def mean_var(model: B200GPE, x):
    """gp.jl:2-5 (vector -> scalars) and :8 (matrix -> vectors)."""
    x = np.asarray(x, float)
    mu, var = model.predict(x)
    return (float(mu[0]), float(var[0])) if x.ndim == 1 else (mu, var)


def myrand(model: B200GPE, x, seed: int = 0, idx_offset: int = 0, joint: bool = True):
    """gp.jl:6-7.  Vector x: an independent posterior sample mu + sigma eps (:6).  Matrix X: ONE joint draw with the full posterior
    covariance (:7 -> EXT rand(gp, X), quirk 9) -- `joint=False` gives independent per-column draws instead (what the batched acquisition
    sweep uses).  eps from the Philox stream keyed by (seed, global index)."""
    x = np.asarray(x, float)
    if x.ndim == 2 and joint:
        return model.rand_joint(x, seed=seed, idx_offset=idx_offset)["sample"]
    r = model.acquire("TS", (), x, seed=seed, idx_offset=idx_offset)
    return float(r["values"][0]) if x.ndim == 1 else r["values"]


def dims(model: B200GPE):
    """gp.jl:9 -- size(model.x) = (D, N)."""
    return model.D, model.nobs


def maxy(model: B200GPE) -> float:
    """gp.jl:10."""
    v = C.c_double()
    check(lib.b200bo_maxy(model._h, C.byref(v)), model._h)
    return v.value


def update(model: B200GPE, x, y):
    """update!(model, x, y), gp.jl:11-18: elastic append; an empty y refits on the current data."""
    y = np.asarray(y, float).ravel()
    if y.size == 0:
        check(lib.b200bo_refit(model._h), model._h)
    else:
        model.append(x, y)


# ------------------------------------------------------------------------------------------------------------
# model optimisers (src/BayesianOptimization.jl:44-49, src/models/gp.jl:20-77)
# ------------------------------------------------------------------------------------------------------------
class ModelOptimizer:
    pass


class NoModelOptimizer(ModelOptimizer):
    """Don't optimize the model ever."""


def defaultoptions_map() -> dict:
    """defaultoptions(::Type{MAPGPOptimizer}) (gp.jl:48-52)."""
    return dict(domean=True, kern=True, noise=True, lik=True, meanbounds=None, kernbounds=None, noisebounds=None,
                likbounds=None, method="LD_LBFGS", maxeval=500)


class MAPGPOptimizer(ModelOptimizer):
    """MAPGPOptimizer(; every = 10, kwargs...) (gp.jl:20-41): MAP hyper-parameter fit every `every` steps."""

    def __init__(self, every: int = 10, **kwargs):
        self.i = 0
        self.every = int(every)
        self.options = {**defaultoptions_map(), **kwargs}


def _bounds(model: B200GPE, o: dict):
    """EXT GP.bounds(gp, noisebounds, meanbounds, kernbounds, likbounds; ...) in theta order; None -> +-Inf."""
    lb, ub = [], []
    nl = model.get_params().size - 1 - (1 if model.mean_kind == "MeanConst" else 0)   # kernel params
    if o["noise"]:
        nbd = o["noisebounds"]
        lb.append(-np.inf if nbd is None else float(nbd[0])); ub.append(np.inf if nbd is None else float(nbd[1]))
    if o["domean"] and model.mean_kind == "MeanConst":
        mb = o["meanbounds"]
        lb.extend([-np.inf] if mb is None else np.ravel(mb[0]).tolist()); ub.extend([np.inf] if mb is None else np.ravel(mb[1]).tolist())
    if o["kern"]:
        kb = o["kernbounds"]
        lb.extend([-np.inf] * nl if kb is None else np.ravel(kb[0]).tolist()); ub.extend([np.inf] * nl if kb is None else np.ravel(kb[1]).tolist())
    return np.array(lb, float), np.array(ub, float)


def _mask_select(model: B200GPE, o: dict):
    sel = []
    n_mean = 1 if model.mean_kind == "MeanConst" else 0
    P = model.num_params
    if o["noise"]:
        sel.append(0)
    if o["domean"]:
        sel.extend(range(1, 1 + n_mean))
    if o["kern"]:
        sel.extend(range(1 + n_mean, P))
    return np.array(sel, int)


def optimizemodel(o, model: B200GPE):
    """optimizemodel!(o, model): NoModelOptimizer -> nothing (BayesianOptimization.jl:49); MAPGPOptimizer ->
    every o.every-th call maximise mll over theta (gp.jl:42-47, 54-77).  The objective/gradient closure (gp.jl:59-64)
    is b200bo_mll_sweep on the device; the outer box-bounded L-BFGS is host glue (NLopt LD_LBFGS in the reference)."""
    if isinstance(o, NoModelOptimizer):
        return None
    ret = None
    if o.i % o.every == 0:
        ret = _map_fit(model, o.options)
    o.i += 1
    return ret


_NLOPT_STATUS = {3: "FTOL_REACHED", 4: "XTOL_REACHED", 5: "MAXEVAL_REACHED", 6: "STALLED", 0: "MAXTIME_REACHED"}


def _map_fit(model: B200GPE, opts: dict):
    """the optimisation of optimizemodel! (gp.jl:54-77) runs inside the library (b200bo_map_fit): objective, gradient and the box-bounded
    L-BFGS iteration; maxeval / ftol / xtol / maxtime are the NLopt options the reference forwards (gp.jl:72)."""
    if model.nobs == 0:
        return None
    sel = _mask_select(model, opts)
    lb, ub = _bounds(model, opts)
    if lb.size != sel.size:
        raise ValueError("bounds do not match the number of optimised parameters")
    theta_full = model.get_params()
    x0 = np.clip(theta_full[sel], lb, ub)
    r = model.map_fit(x0, lb, ub, noise=opts["noise"], domean=opts["domean"], kern=opts["kern"], maxeval=int(opts["maxeval"]),
                      ftol_rel=float(opts.get("ftol_rel", 0.0)), ftol_abs=float(opts.get("ftol_abs", 0.0)), xtol_rel=float(opts.get("xtol_rel", 0.0)),
                      xtol_abs=float(opts.get("xtol_abs", 0.0)), maxtime=float(opts.get("maxtime", 0.0)))
    return float(r["mll"]), r["theta"], _NLOPT_STATUS.get(r["status"], str(r["status"]))
