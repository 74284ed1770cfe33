"""The DIRECT-L search (csrc/direct.h, b200bo_acquire_direct) that stands where the reference calls NLopt :GN_DIRECT_L
(src/acquisition.jl:7-9).  CPU: the product's host state machine against its independent restatement (oracle/direct_oracle.py), point for
point, plus the algorithm's textbook properties.  GPU: the library entry against the restatement driven by the library's own values."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import direct_oracle as dor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host():
    path = os.path.join(ROOT, "oracle", "_build", "libdirect_host.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(path)
    lib.direct_host_create.restype = C.c_void_p
    lib.direct_host_create.argtypes = [C.c_int, C.c_longlong, C.c_int, C.c_int]
    lib.direct_host_destroy.argtypes = [C.c_void_p]
    lib.direct_host_ask.restype = C.c_longlong
    lib.direct_host_ask.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_longlong]
    lib.direct_host_tell.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.direct_host_result.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    return lib


def run_host(fbatch, D, maxeval, width=1, variant=0):
    lib = _host()
    s = lib.direct_host_create(D, maxeval, width, variant)
    buf = np.empty((max(maxeval, 1) + 4 * D, D))
    X, F, batches = [], [], 0
    while True:
        n = lib.direct_host_ask(s, buf.ctypes.data_as(C.POINTER(C.c_double)), buf.shape[0])
        assert n >= 0
        if n == 0:
            break
        P = buf[:n].copy()
        v = np.ascontiguousarray(fbatch(P), float)
        lib.direct_host_tell(s, v.ctypes.data_as(C.POINTER(C.c_double)))
        X.append(P); F.append(v); batches += 1
    bf = C.c_double(); ev = C.c_longlong(); nr = C.c_longlong(); bc = np.empty(D)
    lib.direct_host_result(s, C.byref(bf), bc.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ev), C.byref(nr))
    lib.direct_host_destroy(s)
    return dict(best_f=bf.value, best_c=bc, evals=ev.value, nrect=nr.value, batches=batches, X=np.vstack(X), f=np.concatenate(F))


def neg_branin_unit(P):
    """-branin on [-5,10] x [0,15] (test/branin.jl:1-5) in unit-cube coordinates."""
    x = -5.0 + 15.0 * P[:, 0]; y = 15.0 * P[:, 1]
    return -((y - 5.1 / (4 * np.pi ** 2) * x ** 2 + 5 / np.pi * x - 6) ** 2 + 10 * (1 - 1 / (8 * np.pi)) * np.cos(x) + 10)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("D,width,maxeval", [(2, 1, 400), (6, 1, 1000), (3, 2, 333), (1, 1, 60), (8, 3, 777)])
def test_host_state_machine_equals_restatement(D, width, maxeval, variant):
    rng = np.random.default_rng(D * 100 + width)
    A = rng.normal(size=(4, D)); b = rng.random((4, D))

    def f(P):     # smooth multi-modal, deterministic
        return sum(np.exp(-8 * ((P - b[k]) ** 2).sum(axis=1)) * (1 + 0.3 * k) for k in range(4)) + 0.05 * np.sin(P @ A.T).sum(axis=1)

    h = run_host(f, D, maxeval, width, variant)
    o = dor.direct_l(f, D, maxeval, width, variant)
    assert h["evals"] == o["evals"] == maxeval
    assert h["batches"] == o["batches"] and h["nrect"] == o["nrect"]
    assert np.array_equal(h["X"], o["X"]) and np.array_equal(h["f"], o["f"])
    assert h["best_f"] == o["best_f"] and np.array_equal(h["best_c"], o["best_c"])


def test_first_iterations_follow_direct():
    """iteration 0 = the centre; iteration 1 divides the unit cube along ALL D sides (2 D points at c +- 1/3 e_d);
    the best direction is cut first (its children are the largest)."""
    D = 3
    f = lambda P: -((P - np.array([0.9, 0.5, 0.5])) ** 2).sum(axis=1)
    h = run_host(f, D, 1 + 2 * D)
    assert np.array_equal(h["X"][0], np.full(D, 0.5))
    want = []
    for d in range(D):
        for s in (-1, 1):
            x = np.full(D, 0.5); x[d] += s * (1.0 / 3.0); want.append(x)
    assert np.allclose(h["X"][1:], np.array(want), atol=1e-16)
    assert h["best_f"] == f(h["X"]).max() and h["nrect"] == 1 + 2 * D


def test_original_direct_cuts_non_cubes_along_one_side_only():
    """NLopt GN_DIRECT (cdirect which_div = 0): the unit cube is trisected along all D sides; its best child -- no longer a cube -- along ONE
    longest side only, and every rectangle tied for the best value of a hull class divides"""
    D = 3
    f = lambda P: -((P - np.array([0.9, 0.5, 0.5])) ** 2).sum(axis=1)
    h = run_host(f, D, 1 + 2 * D + 40, variant=1)
    assert np.array_equal(h["X"][0], np.full(D, 0.5)) and h["batches"] >= 3
    second = h["X"][1 + 2 * D:]                       # iteration 2: each dividing rectangle contributes exactly one +- pair along one axis
    moved = np.abs(second[0::2] - second[1::2]) > 0
    assert np.all(moved.sum(axis=1) == 1)
    lb = run_host(neg_branin_unit, 2, 2000, variant=1)
    assert -lb["best_f"] - 0.397887 < 1e-3


def test_maxeval_is_exact_and_nan_never_wins():
    calls = []

    def f(P):
        calls.append(len(P))
        v = -((P - 0.3) ** 2).sum(axis=1)
        v[P[:, 0] > 0.6] = np.nan
        return v

    for me in (1, 2, 7, 50, 51):
        calls.clear()
        h = run_host(f, 2, me)
        assert h["evals"] == me == sum(calls)
        assert np.isfinite(h["best_f"]) and h["best_c"][0] <= 0.6
    allnan = run_host(lambda P: np.full(len(P), np.nan), 2, 30)
    assert allnan["best_f"] == -np.inf and allnan["evals"] == 30


def test_branin_global_optimum_within_reference_budget():
    """the reference's default budget for this method is maxeval = 2000 (src/acquisition.jl:8)"""
    h = run_host(neg_branin_unit, 2, 2000)
    assert -h["best_f"] - 0.397887 < 1e-4
    assert h["batches"] < 250          # ~10 evaluations per launch even at width 1
    w = run_host(neg_branin_unit, 2, 2000, width=4)
    assert -w["best_f"] - 0.397887 < 1e-3 and w["batches"] < h["batches"]


@pytest.mark.gpu
@pytest.mark.parametrize("kind,variant", [("EI", 0), ("TS", 0), ("MaxMean", 0), ("EI", 1)])
def test_library_direct_equals_restatement_on_library_values(kind, variant):
    import b200bo
    from oracle import gp_oracle as orc
    rng = np.random.default_rng(5)
    D, N = 4, 300
    X = rng.random((D, N)); y = np.sin(3 * X.sum(axis=0)) + 0.1 * rng.normal(size=N)
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.SEArd(np.full(D, -0.7), 0.0), logNoise=-2.0, capacity=N)
    g.fit(X, y)
    lb = np.array([-0.5, 0.0, 0.2, 0.0]); ub = np.array([1.0, 1.0, 0.9, 2.0])
    params = (float(y.max()),) if kind == "EI" else ()
    me, seed = 500, 11
    r = g.acquire_direct(kind, params, lb, ub, maxeval=me, seed=seed, want_trace=True, variant=variant)
    assert r["evals"] == me and 0 <= r["best_index"] < me
    assert r["values"][r["best_index"]] == r["best_value"] and np.array_equal(r["X"][:, r["best_index"]], r["best_x"])
    assert r["best_index"] == int(np.argmax(np.where(np.isnan(r["values"]), -np.inf, r["values"])))      # first strict maximum
    state = dict(n=0)

    def f(P):      # the library's own values at the same global evaluation indices
        Xs = lb[:, None] + P.T * (ub - lb)[:, None]
        v = g.acquire(kind, params, Xs, seed=seed, idx_offset=state["n"])["values"]
        state["n"] += len(P)
        return v

    o = dor.direct_l(f, D, me, 1, variant)
    Xo = lb[None, :] + o["X"] * (ub - lb)[None, :]
    assert np.array_equal(r["X"].T, Xo) and np.array_equal(r["values"], o["f"])
    assert r["best_value"] == o["best_f"] and r["batches"] == o["batches"]
    assert np.array_equal(r["best_x"], lb + o["best_c"] * (ub - lb))
    # and the values are the reference's functor on the oracle posterior (north_star tolerance)
    oc = orc.GPOracle(D, "SEArd", "MeanConst", ll=np.full(D, -0.7), lsigma=0.0, lognoise=-2.0, beta=0.0).fit(X, y)
    mu, s2 = oc.predict(r["X"])
    eps = orc.philox_normal(seed, np.arange(me)) if kind == "TS" else None      # evaluation e draws from the stream at global index e
    a = orc.acq_value(kind, params, mu, s2, eps)
    assert np.all(np.abs(r["values"] - a) <= 1e-5 * np.abs(a) + 1e-12)


@pytest.mark.gpu
def test_direct_finds_the_posterior_mode_of_the_one_point_gp():
    """test/acquisition.jl:11-12: the arg-max of the posterior mean of the 1-point GP is the observation itself"""
    import b200bo
    g = b200bo.B200GPE.from_data(np.array([[1.0]]), np.array([2.0]), kernel=b200bo.SEIso(0.0, 0.0), mean=b200bo.MeanZero(), logNoise=-2.0)
    r = g.acquire_direct("MaxMean", (), np.array([-3.0]), np.array([5.0]), maxeval=200)
    assert abs(r["best_x"][0] - 1.0) < 1e-3 and r["evals"] == 200


@pytest.mark.gpu
def test_acquire_max_with_the_reference_thompson_defaults():
    """defaultoptions(GPE, ThompsonSamplingSimple) = (method = :GN_DIRECT_L, restarts = 1, maxeval = 2000) (src/acquisition.jl:7-9):
    one start, so the answer comes from the DIRECT-L search"""
    import b200bo as bo
    rng = np.random.default_rng(1)
    X = rng.random((2, 40)) * np.array([[15.0], [15.0]]) + np.array([[-5.0], [0.0]])
    y = -np.array([orc_branin(X[:, i]) for i in range(40)])
    model = bo.ElasticGPE(2, mean=bo.MeanConst(-10.0), kernel=bo.SEArd([1.0, 1.0], 4.0), logNoise=-2.0, capacity=64)
    model.append(X, y)
    ac = bo.ThompsonSamplingSimple()
    opt = bo.nlopt_setup(ac, model, [-5.0, 0.0], [10.0, 15.0], dict(method="GN_DIRECT_L", restarts=1, maxeval=2000))
    assert not opt.gradient
    f, x = bo.acquire_max(opt, [-5.0, 0.0], [10.0, 15.0], 1)
    assert np.isfinite(f) and np.all(x >= [-5.0, 0.0]) and np.all(x <= [10.0, 15.0])
    one = model.acquire("TS", (), np.array([[2.5], [7.5]]), seed=opt.seed - 1)      # the single LHS start of the sweep
    assert f >= one["best_value"]
    mm = bo.acquire_max(bo.MaxMean(), model, [-5.0, 0.0], [10.0, 15.0], dict(method="GN_DIRECT_L", restarts=1, maxeval=2000))
    mu, _ = model.predict(np.asarray(mm[1]).reshape(2, 1))
    assert mm[0] == mu[0]
    grid = model.predict(np.stack(np.meshgrid(np.linspace(-5, 10, 61), np.linspace(0, 15, 61)), 0).reshape(2, -1))[0]
    assert mm[0] >= grid.max() - 1e-3 * abs(grid.max())


def orc_branin(x):
    from oracle import gp_oracle as orc
    return float(orc.branin(x[0], x[1]))


def test_hull_selection_matches_the_definition_of_potentially_optimal():
    """Jones' definition (epsilon = 0, maximisation): size class j is potentially optimal iff some rate K >= 0 has
    f_j + K d_j >= f_i + K d_i for every class i.  The hull routine of the restatement (and, through the point-for-point tests above, of
    csrc/direct.h) must select exactly the classes the definition admits among those at least as large as the incumbent's."""
    rng = np.random.default_rng(3)
    for trial in range(300):
        n = int(rng.integers(1, 9))
        levels = sorted(rng.choice(np.arange(0, 12), size=n, replace=False))          # distinct size classes, s ascending = size descending
        d = np.array([3.0 ** -int(s) for s in levels])
        f = np.round(rng.normal(size=n), 1 if trial % 2 else 3)                        # coarse values: ties and collinear triples occur
        star = int(np.argmax(f))                                                       # first maximum = largest size among ties
        pts = [(d[i], f[i], levels[i]) for i in range(star, -1, -1)]                   # from the incumbent's class towards larger sizes
        got = set(dor._upper_right_hull(pts))
        want = set()
        for j in range(star + 1):
            lo, hi = 0.0, np.inf                                                       # feasible K: f_j - f_i >= K (d_i - d_j) for all i
            ok = True
            for i in range(n):
                if i == j:
                    continue
                dd, df = d[i] - d[j], f[j] - f[i]
                if dd > 0:
                    hi = min(hi, df / dd)
                elif dd < 0:
                    lo = max(lo, df / dd)
                elif df < 0:
                    ok = False
            if ok and lo <= hi and hi >= 0:
                want.add(levels[j])
        assert got == want, (levels, f.tolist(), got, want)
