"""ctypes loader for oracle/_build/liboracle.so (the C restatement; TEST INFRASTRUCTURE ONLY -- see oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import gp_oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_FAM = {"SE": 0, "Mat12": 1, "Mat32": 2, "Mat52": 3}
_dp = C.POINTER(C.c_double)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def load():
    if not os.path.exists(_SO):
        build()
    lib = C.CDLL(_SO)
    lib.orc_fit.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]
    lib.orc_fit.restype = C.c_int
    lib.orc_acquire.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, _dp, _dp, C.c_int, C.c_double,
                                C.c_double, C.c_uint64, C.c_int64, _dp, C.c_int64, C.c_int, _dp, _dp, _dp, _dp, _dp,
                                C.POINTER(C.c_int64), C.c_int]
    lib.orc_acquire.restype = C.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def fam_of(kernel: str) -> int:
    return _FAM[kernel[:-3]]


class COracle:
    """Same model state as gp_oracle.GPOracle, arithmetic in plain C (+OpenMP over candidates)."""

    def __init__(self, gp: orc.GPOracle, refit_in_c: bool = False):
        self.lib = load()
        self.gp = gp
        self.D, self.N = gp.D, gp.y.size
        self.X = np.asfortranarray(gp.X)
        self.inv_ell = np.ascontiguousarray(gp.inv_ell)
        self.fam = fam_of(gp.kernel)
        self.beta = gp.beta if gp.mean == "MeanConst" else 0.0
        if refit_in_c:
            self.U = np.zeros((self.N, self.N), order="F")
            self.alpha = np.zeros(self.N)
            mll = C.c_double()
            noise = float(np.exp(2 * gp.lognoise) + orc.EPS)
            rc = self.lib.orc_fit(self.D, self.N, self.fam, _p(self.X), _p(np.ascontiguousarray(gp.y)), _p(self.inv_ell), gp.sf2,
                                  noise, self.beta, _p(self.U), _p(self.alpha), C.byref(mll))
            if rc != 0:
                raise np.linalg.LinAlgError(f"not positive definite at {rc - 1}")
            self.mll = mll.value
        else:
            self.U = np.asfortranarray(gp.U)
            self.alpha = np.ascontiguousarray(gp.alpha)
            self.mll = gp.mll

    def acquire(self, kind, params, Xs, seed=0, idx_offset=0, want_grad=False, nthreads=0):
        Xs = np.asfortranarray(np.asarray(Xs, float).reshape(self.D, -1))
        M = Xs.shape[1]
        acq = -1 if kind is None else orc.ACQS.index(kind)
        p = list(params) + [0.0, 0.0]
        vals, mu, var = np.empty(M), np.empty(M), np.empty(M)
        grad = np.empty((self.D, M), order="F") if want_grad else None
        bv, bi = C.c_double(), C.c_int64()
        used = self.lib.orc_acquire(self.D, self.N, self.fam, _p(self.X), _p(self.inv_ell), self.gp.sf2, self.beta, _p(self.U),
                                    _p(self.alpha), acq, p[0], p[1], seed, idx_offset, _p(Xs), M, int(want_grad), _p(vals), _p(mu),
                                    _p(var), _p(grad), C.byref(bv), C.byref(bi), nthreads)
        return dict(values=vals, mu=mu, var=var, grad=grad, best_value=bv.value, best_index=bi.value, threads=used)
