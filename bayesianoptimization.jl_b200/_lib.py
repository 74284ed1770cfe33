"""ctypes binding of libb200bo.so (include/b200bo.h) -- byte-for-byte the calls a Julia `ccall` host makes.

There is NO fallback: if the shared library is missing this module raises at import; if there is no B200 every
compute entry returns B200BO_ERR_CUDA and `check` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200bo.so")

OK, ERR_ARG, ERR_CUDA, ERR_NOTPD, ERR_STATE, ERR_ALLOC, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6

KERNEL_KINDS = {"SEIso": 0, "SEArd": 1, "Mat12Iso": 2, "Mat12Ard": 3, "Mat32Iso": 4, "Mat32Ard": 5, "Mat52Iso": 6,
                "Mat52Ard": 7}
MEAN_KINDS = {"MeanZero": 0, "MeanConst": 1}
ACQ_KINDS = {"PI": 0, "EI": 1, "UCB": 2, "TS": 3, "MI": 4, "MaxMean": 5}
MASK_NOISE, MASK_MEAN, MASK_KERN = 1, 2, 4
T_KMAT, T_CHOL, T_SYRK, T_ALPHA, T_ACQ, T_MLL, T_ACQ_GEMM = range(7)


class Best(C.Structure):
    _fields_ = [("value", C.c_double), ("index", C.c_int64)]


class B200BOError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200bo error {code}: {msg}")
        self.code = code


_dp = C.POINTER(C.c_double)
_H = C.c_void_p

# name -> argtypes (restype is int32 unless listed in _RESTYPES); mirrors include/b200bo.h one to one
PROTOTYPES = {
    "b200bo_create": [C.POINTER(_H), C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32],
    "b200bo_destroy": [_H],
    "b200bo_create_multi": [C.POINTER(_H), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int64, C.c_int32, C.c_int32],
    "b200bo_num_gpus": [_H, C.POINTER(C.c_int32)],
    "b200bo_comm_unique_id": [C.c_char_p],
    "b200bo_comm_init_rank": [_H, C.c_int32, C.c_int32, C.c_char_p],
    "b200bo_comm_destroy": [_H],
    "b200bo_last_error": [_H],
    "b200bo_set_stream": [_H, C.c_void_p],
    "b200bo_sync": [_H],
    "b200bo_num_params": [_H, C.POINTER(C.c_int32)],
    "b200bo_set_params": [_H, _dp, C.c_int32],
    "b200bo_get_params": [_H, _dp, C.c_int32],
    "b200bo_set_priors": [_H, C.c_int32, C.POINTER(C.c_int32), _dp, _dp],
    "b200bo_fit": [_H, _dp, _dp, C.c_int64],
    "b200bo_append": [_H, _dp, _dp, C.c_int64],
    "b200bo_refit": [_H],
    "b200bo_dims": [_H, C.POINTER(C.c_int32), C.POINTER(C.c_int64)],
    "b200bo_maxy": [_H, _dp],
    "b200bo_get_data": [_H, _dp, _dp],
    "b200bo_get_mll": [_H, _dp],
    "b200bo_get_alpha": [_H, _dp],
    "b200bo_get_factor": [_H, _dp],
    "b200bo_jitter_tries": [_H, C.POINTER(C.c_int32)],
    "b200bo_predict": [_H, _dp, C.c_int64, _dp, _dp],
    "b200bo_rand_joint": [_H, _dp, C.c_int64, C.c_uint64, C.c_int64, _dp, _dp, C.POINTER(C.c_int32)],
    "b200bo_acquire": [_H, C.c_int32, _dp, C.c_int32, _dp, C.c_int64, C.c_uint64, C.c_int64, _dp, _dp, _dp, _dp,
                       C.POINTER(Best), _dp],
    "b200bo_mll_sweep": [_H, _dp, C.c_int32, C.c_int32, C.c_int32, _dp, _dp],
    "b200bo_kmat_dev": [_H, C.c_void_p, C.c_int64],
    "b200bo_predict_dev": [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p],
    "b200bo_acquire_dev": [_H, C.c_int32, _dp, C.c_int32, C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "b200bo_lhs": [_H, _dp, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, _dp],
    "b200bo_acquire_lhs": [_H, C.c_int32, _dp, C.c_int32, _dp, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, _dp,
                           C.POINTER(Best), _dp],
    "b200bo_acquire_ascent": [_H, C.c_int32, _dp, C.c_int32, _dp, C.c_int64, _dp, _dp, C.c_int32, C.c_double, C.c_int64, _dp, _dp,
                              C.POINTER(Best), _dp],
    "b200bo_sobol": [_H, _dp, _dp, C.c_uint64, C.c_int64, _dp],
    "b200bo_acquire_lbfgs": [_H, C.c_int32, _dp, C.c_int32, _dp, C.c_int64, _dp, _dp, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double,
                             C.c_double, C.c_double, C.c_int64, _dp, _dp, _dp, C.POINTER(Best), _dp],
    "b200bo_acquire_direct": [_H, C.c_int32, _dp, C.c_int32, _dp, _dp, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_uint64, _dp, _dp,
                              C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(Best), _dp],
    "b200bo_map_fit": [_H, _dp, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                       _dp, _dp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "b200bo_kmat": [_H, _dp],
    "b200bo_last_timing_ms": [_H, C.c_int32, C.POINTER(C.c_float)],
    "b200bo_launch_count": [_H, C.POINTER(C.c_int64)],
    "b200bo_fp64_peak_tflops": [_H, _dp],
    "b200bo_set_syrk_engine": [_H, C.c_int32],
    "b200bo_set_acq_engine": [_H, C.c_int32],
    "b200bo_i8_peak_tops": [_H, _dp],
    "b200bo_set_knob": [_H, C.c_char_p, C.c_int64],
    "b200bo_version": [],
}
_RESTYPES = {"b200bo_last_error": C.c_char_p}


def load(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(libb200bo has no CPU fallback)")
    lib = C.CDLL(path)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library drift
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int32)
    return lib


lib = load()


def check(rc: int, h=None):
    if rc != OK:
        msg = lib.b200bo_last_error(h)
        raise B200BOError(rc, msg.decode() if msg else "?")


def dptr(a: np.ndarray):
    """pointer to a float64 numpy array (C- or F-contiguous: the caller owns the layout)"""
    if a is None:
        return None
    assert a.dtype == np.float64 and (a.flags.c_contiguous or a.flags.f_contiguous)
    return a.ctypes.data_as(_dp)
