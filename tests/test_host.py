"""CPU tests of the host mirror and the C-ABI boundary (no compute calls: there is no GPU here)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import b200bo
from b200bo import _lib
from b200bo import acquisitionfunctions as af
from oracle import gp_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200bo.h")).read()
    declared = set(re.findall(r"B200BO_API\s+(?:int32_t|const char\*)\s+(b200bo_\w+)\s*\(", hdr))
    assert len(declared) >= 28
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200bo.h but not exported by libb200bo.so"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes drifted from the header"
    assert _lib.lib.b200bo_version() == 100


def test_enums_match_header():
    hdr = open(os.path.join(ROOT, "include", "b200bo.h")).read()
    vals = {k: int(v) for k, v in re.findall(r"(B200BO_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)}
    for name, v in _lib.KERNEL_KINDS.items():
        assert vals["B200BO_KERNEL_" + name.upper()] == v
    for name, v in _lib.ACQ_KINDS.items():
        assert vals["B200BO_ACQ_" + name.upper()] == v
    assert (vals["B200BO_ERR_ARG"], vals["B200BO_ERR_CUDA"], vals["B200BO_ERR_NOTPD"]) == (_lib.ERR_ARG, _lib.ERR_CUDA, _lib.ERR_NOTPD)
    assert [orc.ACQS.index(k) for k in ("PI", "EI", "UCB", "TS", "MI", "MaxMean")] == [_lib.ACQ_KINDS[k] for k in
                                                                                  ("PI", "EI", "UCB", "TS", "MI", "MaxMean")]


def test_no_cpu_fallback_and_argument_errors():
    import torch
    h = ctypes.c_void_p()
    assert _lib.lib.b200bo_create(ctypes.byref(h), 0, 0, 10, 1, 1) == _lib.ERR_ARG            # D out of range
    assert _lib.lib.b200bo_create(ctypes.byref(h), 0, 2, 10, 99, 1) == _lib.ERR_ARG           # unknown kernel
    assert _lib.lib.b200bo_create(None, 0, 2, 10, 1, 1) == _lib.ERR_ARG
    assert b"null handle" in _lib.lib.b200bo_last_error(None)
    if not torch.cuda.is_available():
        assert _lib.lib.b200bo_create(ctypes.byref(h), 0, 2, 10, 1, 1) == _lib.ERR_CUDA       # fails loudly, no fallback
        with pytest.raises(_lib.B200BOError):
            b200bo.B200GPE(2)
    assert _lib.lib.b200bo_fit(None, None, None, 0) == _lib.ERR_ARG
    assert _lib.lib.b200bo_destroy(None) == _lib.OK


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bayesianoptimization.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|liboracle|#include\s+\"[^\"]*oracle", src, flags=re.M), f


def test_functor_scalars_match_oracle():
    for mu, s2 in [(0.3, 0.01), (1.2, 0.5), (-0.4, 2.0), (0.7, 0.0), (0.2, 0.0)]:
        assert af.ExpectedImprovement(0.5)(mu, s2) == pytest.approx(float(orc.acq_value("EI", (0.5,), mu, s2)), rel=1e-14)
        assert af.ProbabilityOfImprovement(0.5)(mu, s2) == pytest.approx(float(orc.acq_value("PI", (0.5,), mu, s2)), rel=1e-14)
        assert af.UpperConfidenceBound(beta_t=2.0)(mu, s2) == pytest.approx(float(orc.acq_value("UCB", (2.0,), mu, s2)))
        mi = af.MutualInformation(alpha=2.25, gamma_hat=0.2)
        assert mi.params() == (1.5, 0.2)
        assert mi(mu, s2) == pytest.approx(float(orc.acq_value("MI", (1.5, 0.2), mu, s2)))
    assert af.ExpectedImprovement().tau == -math.inf and af.ProbabilityOfImprovement().tau == -math.inf
    assert [c.kind for c in (af.ProbabilityOfImprovement, af.ExpectedImprovement, af.UpperConfidenceBound, af.ThompsonSamplingSimple,
                             af.MutualInformation, af.MaxMean)] == ["PI", "EI", "UCB", "TS", "MI", "MaxMean"]


def test_defaultoptions_and_search_options_passthrough():
    # acquisition.jl:4-9 keys; test/acquisition.jl:7-9 pass-through (maxeval / maxtime / ftol_abs)
    o = b200bo.defaultoptions(b200bo.B200GPE, af.ExpectedImprovement)
    assert o["method"] == "LD_LBFGS" and o["maxeval"] == 2000 and o["restarts"] >= 10
    t = b200bo.defaultoptions(b200bo.B200GPE, af.ThompsonSamplingSimple)
    assert t["method"] == "GN_DIRECT_L" and t["maxeval"] == 2000

    class FakeModel:            # nlopt_setup only needs setparams!(a, model) to work
        pass
    opt = b200bo.nlopt_setup(af.MaxMean(), FakeModel(), [-5.0], [5.0], {**o, "maxtime": 3.0, "ftol_abs": np.finfo(float).eps})
    assert opt.maxeval == 2000 and opt.maxtime == 3.0 and opt.ftol_abs == np.finfo(float).eps
    assert opt.gradient is True
    assert b200bo.nlopt_setup(af.ThompsonSamplingSimple(), FakeModel(), [0.0], [1.0], t).gradient is False


def test_merge_with_defaults_errors():
    # test/BayesianOptimization.jl:1-24
    f = lambda x: 0.0
    with pytest.raises(ValueError, match="unsupported keyword"):
        b200bo.merge_with_defaults(f, [0.0], [1.0], dict(bogus=1))
    with pytest.raises(ValueError, match="length of lowerbounds"):
        b200bo.merge_with_defaults(f, [0.0, 0.0], [1.0], {})
    sentinel = object()
    args, kwargs = b200bo.merge_with_defaults(f, [0.0], [1.0], dict(model=sentinel, acquisition="a", modeloptimizer="m",
                                                                    sense=b200bo.Min, maxiterations=7))
    assert args[0] is f and args[1] is sentinel and args[2] == "a" and args[3] == "m" and args[4] == [0.0] and args[5] == [1.0]
    assert kwargs == dict(sense=b200bo.Min, maxiterations=7)


def test_counters():
    # test/utils.jl
    from b200bo.utils import IterationCounter, DurationCounter, isdone, step, init
    it = IterationCounter(0, 0, 2)
    assert not isdone(it); step(it); step(it); assert isdone(it) and it.i == 2
    init(it); assert it.c == 0 and it.i == 2
    d = DurationCounter(0.0, 0.05, 0.0, 0.0)
    init(d); assert not isdone(d)
    import time; time.sleep(0.06); assert isdone(d)


def test_lhs_and_sobol_iterators():
    rng = np.random.default_rng(1)
    S = b200bo.latin_hypercube_sampling([-5.0, 0.0], [10.0, 15.0], 32, rng)
    assert S.shape == (2, 32) and S.flags.f_contiguous
    for d, (lo, hi) in enumerate([(-5.0, 10.0), (0.0, 15.0)]):
        assert sorted(np.floor((S[d] - lo) / ((hi - lo) / 32)).astype(int)) == list(range(32))
    it = b200bo.ScaledLHSIterator([0.0], [1.0], 5, rng)
    assert len(it) == 5 and len(list(it)) == 5
    sob = b200bo.ScaledSobolIterator([-5.0, 0.0], [10.0, 15.0], 10)
    pts = np.array(list(sob))
    assert len(sob) == 10 and pts.shape == (10, 2)
    assert np.all(pts >= [-5.0, 0.0]) and np.all(pts <= [10.0, 15.0]) and len({tuple(p) for p in pts}) == 10
    assert len(b200bo.ScaledSobolIterator([0.0], [1.0], 0)) == 0
    # Sobol.jl's published 2-D sequence (its README): .5 .5 | .75 .25 | .25 .75 | .375 .375 | .875 .875 | .625 .125 | .125 .625 | ...
    # skip(seq, 3) is NOT exact: it advances by 2^floor(log2(3+1)) = 4 points, so the iterator starts at the 5th (src/utils.jl:80)
    s3 = b200bo.ScaledSobolIterator([0.0, 0.0], [1.0, 1.0], 3)
    assert np.array_equal(np.array(list(s3)), [[0.875, 0.875], [0.625, 0.125], [0.125, 0.625]])
    assert np.array_equal(np.array(list(s3))[0], [0.1875, 0.3125])       # a second pass continues the sequence (next! on it.seq)
    s10 = b200bo.ScaledSobolIterator([0.0, 0.0], [1.0, 1.0], 10)         # N = 10 = 5 D: skips 8, not 10
    assert np.array_equal(next(iter(s10)), [0.6875, 0.8125])
    with pytest.raises(ValueError):
        b200bo.latin_hypercube_sampling([0.0, 1.0], [1.0], 4)


def test_shard_bounds_and_select_best():
    from b200bo.dist import shard_bounds, select_best
    for M, ws in [(10, 3), (1048576, 8), (5, 8), (0, 2)]:
        blocks = [shard_bounds(M, ws, r) for r in range(ws)]
        assert blocks[0][0] == 0 and blocks[-1][1] == M
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(ws - 1))
        assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
    assert select_best([1.0, 3.0, 3.0], [5, 9, 7]) == (3.0, 7)          # tie -> lowest global index
    assert select_best([np.nan, 2.0], [0, 4]) == (2.0, 4)               # NaN never wins
    assert select_best([-np.inf, -np.inf], [-1, -1]) == (-np.inf, -1)   # nothing beat -Inf anywhere
    assert select_best([5.0, 1.0], [-1, 3]) == (1.0, 3)                 # a rank with no winner is ignored
