"""GPU parity tests: the CUDA path (through the C ABI, libb200bo.so) against the CPU restatement and the golden
fixtures.  Tolerances are the north_star's: posterior mean/var <= 1e-6 rel, acquisition values <= 1e-5 rel,
selected index bit-exact for deterministic acquisitions.  Relative checks carry an absolute floor ATOL because
the reference's own formulae (0.5(1+erf) etc., src/utils.jl:48-49) cancel to ~1e-16 in the tails.
"""
import glob
import os

import numpy as np
import pytest

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu

RTOL_POST, RTOL_ACQ, ATOL = 1e-6, 1e-5, 1e-12


@pytest.fixture(scope="module")
def bo():
    import b200bo
    return b200bo


def close(a, b, rtol, atol=ATOL):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return bool(np.all(np.abs(a - b) <= rtol * np.abs(b) + atol))


def relmax(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def make_pair(bo, kern, mean, D, N, seed, lognoise=-2.0, capacity=None):
    rng = np.random.default_rng(seed)
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    ll = rng.normal(np.log(np.sqrt(D) * 0.3), 0.1, 1 if kern.endswith("Iso") else D)
    beta = 0.3 if mean == "MeanConst" else 0.0
    o = orc.GPOracle(D, kern, mean, ll=ll, lsigma=0.1, lognoise=lognoise, beta=beta).fit(X, y)
    g = bo.B200GPE(D, mean=bo.MeanConst(beta) if mean == "MeanConst" else bo.MeanZero(), kernel=bo.gp._Kernel(kern, ll, 0.1),
                   logNoise=lognoise, capacity=capacity or N)
    g.fit(X, y)
    return rng, o, g, X, y


ACQS = lambda D, N, y: [("EI", (float(np.quantile(y, 0.9)),)), ("PI", (float(np.quantile(y, 0.9)),)),
                        ("UCB", (orc.brochu_beta(D, N),)), ("MI", (1.0, 0.25)), ("MaxMean", ())]


@pytest.mark.parametrize("name", [os.path.basename(f)[:-4] for f in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))])
def test_golden_fixture(bo, golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    D, kern, mean = int(z["D"]), str(z["kernel"]), str(z["mean"])
    th = z["theta"]
    nm = 1 if mean == "MeanConst" else 0
    g = bo.B200GPE(D, mean=bo.MeanConst(th[1]) if nm else bo.MeanZero(), kernel=bo.gp._Kernel(kern, th[1 + nm:-1], th[-1]),
                   logNoise=th[0], capacity=z["y"].size)
    g.fit(z["X"], z["y"])
    assert np.array_equal(g.get_params(), th)
    assert close(g.alpha, z["alpha"], 1e-8, 1e-10 * np.abs(z["alpha"]).max())
    assert abs(g.mll - float(z["mll"])) <= 1e-10 * abs(float(z["mll"]))
    U = g.factor
    assert close(np.diag(U), z["Udiag"], 1e-10) and close(U[:, -1], z["Ucol_last"], 1e-9, 1e-12)
    mu, var = g.predict(z["Xs"])
    assert close(mu, z["mu"], RTOL_POST) and close(var, z["var"], RTOL_POST)
    for k in ("EI", "PI", "UCB", "MI", "MaxMean"):
        r = g.acquire(k, z[f"{k}_params"], z["Xs"], want_grad=True)
        assert close(r["values"], z[f"{k}_values"], RTOL_ACQ), k
        assert r["best_index"] == int(z[f"{k}_best"]), k                       # bit-exact selected index
        assert relmax(r["grad"], z[f"{k}_grad"]) < 1e-8, k
        assert np.array_equal(r["best_x"], z["Xs"][:, r["best_index"]])
    r = g.acquire("TS", (), z["Xs"], seed=int(z["TS_seed"]), idx_offset=int(z["TS_offset"]))
    assert close(r["values"], z["TS_values"], RTOL_ACQ)
    mll, dmll = g.mll_sweep(np.stack([th, z["theta2"]], axis=1))
    assert abs(mll[0] - float(z["mll"])) <= 1e-10 * abs(float(z["mll"])) and abs(mll[1] - float(z["mll2"])) <= 1e-10 * abs(float(z["mll2"]))
    assert relmax(dmll[:, 0], z["dmll"]) < 1e-8 and relmax(dmll[:, 1], z["dmll2"]) < 1e-8
    assert np.array_equal(g.get_params(), th)                                  # sweep leaves the model untouched
    assert close(g.predict(z["Xs"])[0], z["mu"], RTOL_POST)                    # and the factor is restored lazily


@pytest.mark.parametrize("kern,mean,D,N,M", [("SEArd", "MeanConst", 3, 40, 10), ("SEIso", "MeanZero", 3, 4, 2),
                                             ("Mat52Ard", "MeanConst", 6, 300, 200), ("Mat32Iso", "MeanZero", 4, 129, 65),
                                             ("Mat12Ard", "MeanConst", 5, 257, 130), ("Mat52Iso", "MeanConst", 2, 128, 64),
                                             ("SEArd", "MeanConst", 8, 1000, 1000), ("SEArd", "MeanConst", 32, 640, 300)])
def test_against_oracle(bo, kern, mean, D, N, M):
    rng, o, g, X, y = make_pair(bo, kern, mean, D, N, seed=D * 1000 + N)
    K = g.kmat()
    S = o.cov(o.X, o.X); S[np.diag_indices(N)] += np.exp(2 * o.lognoise) + orc.EPS
    assert relmax(K, S) < 1e-14 and np.array_equal(K, K.T)
    assert relmax(g.factor, o.U) < 1e-11
    assert relmax(g.alpha, o.alpha) < 1e-9 and abs(g.mll - o.mll) < 1e-11 * abs(o.mll)
    Xs = rng.random((D, M)); Xs[:, 0] = X[:, 3]
    mu, var = g.predict(Xs)
    mo, vo = o.predict(Xs)
    assert close(mu, mo, RTOL_POST) and close(var, vo, RTOL_POST)
    for kind, par in ACQS(D, N, y):
        r = g.acquire(kind, par, Xs, want_grad=True, want_mu_var=True)
        a, gr = orc.acq_grad(o, kind, par, Xs)
        assert close(r["values"], a, RTOL_ACQ), kind
        assert r["best_index"] == orc.first_strict_argmax_np(a), kind
        assert relmax(r["grad"], gr) < 1e-8, kind
        assert np.array_equal(r["mu"], mu) and np.array_equal(r["var"], var)   # same launch path, same bits
    r = g.acquire("TS", (), Xs, seed=50, idx_offset=1000)
    ts = orc.acq_value("TS", (), mo, vo, eps=orc.philox_normal(50, 1000 + np.arange(M)))
    assert close(r["values"], ts, RTOL_ACQ) and r["best_index"] - 1000 == orc.first_strict_argmax_np(ts)


@pytest.mark.parametrize("kern,D,N", [("SEArd", 5, 300), ("Mat52Ard", 3, 700), ("SEIso", 4, 1300), ("Mat32Ard", 6, 2048)])
def test_mll_gradient_block_inversion(bo, kern, D, N):
    """Sigma^-1 by recursive block inversion (3, 6, 11 and 16 panels: uneven merges) against the oracle's mll gradient."""
    rng, o, g, X, y = make_pair(bo, kern, "MeanConst", D, N, seed=7 * N + D)
    th = g.get_params()
    th2 = th.copy(); th2[0] -= 0.4; th2[2:] += 0.15
    mll, dmll = g.mll_sweep(np.stack([th, th2], axis=1))
    for k, t in enumerate((th, th2)):
        mo, go = o.mll_dmll(t)
        assert abs(mll[k] - mo) <= 1e-10 * abs(mo)
        assert relmax(dmll[:, k], go) < 1e-8
    assert np.array_equal(g.get_params(), th)


def test_kmat_extreme_signal_variance(bo):
    """sf2 outside [1e-12, 1e12] is applied after the table-based exponential (K1's `post` path); identical structure otherwise."""
    rng = np.random.default_rng(11)
    D, N = 5, 200
    X = rng.random((D, N)); y = rng.standard_normal(N)
    for ls in (15.0, -15.0):
        ll = np.full(D, -0.5)
        o = orc.GPOracle(D, "SEArd", "MeanZero", ll=ll, lsigma=ls, lognoise=ls - 2.0, beta=0.0).fit(X, y)
        g = bo.B200GPE(D, mean=bo.MeanZero(), kernel=bo.SEArd(ll, ls), logNoise=ls - 2.0, capacity=N)
        g.fit(X, y)
        S = o.cov(o.X, o.X); S[np.diag_indices(N)] += np.exp(2 * o.lognoise) + orc.EPS
        K = g.kmat()
        assert relmax(K, S) < 1e-13 and np.array_equal(K, K.T)
        assert relmax(g.alpha, o.alpha) < 1e-8


@pytest.mark.parametrize("N,D", [(1100, 4), (2560, 8), (4096, 16)])
def test_trailing_update_engines_agree(bo, N, D):
    """K = 512 trailing updates on tcgen05 (int8 slices; 64-wide tiles = default, 128-wide two-pass variant) and on DMMA give the
    same factor to FP64 round-off, and all match LAPACK; the tcgen05 path is deterministic."""
    rng = np.random.default_rng(N)
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    ll = np.full(D, np.log(np.sqrt(D) * 0.3))
    o = orc.GPOracle(D, "SEArd", "MeanConst", ll=ll, lsigma=0.1, lognoise=-2.0, beta=0.2).fit(X, y)
    res = {}
    for eng in (1, 0, 2, 1):
        g = bo.B200GPE(D, mean=bo.MeanConst(0.2), kernel=bo.SEArd(ll, 0.1), logNoise=-2.0, capacity=N)
        g.set_syrk_engine(eng)
        g.fit(X, y)
        assert g.jitter_tries == 0
        U = g.factor
        assert relmax(U, o.U) < 1e-11 and relmax(g.alpha, o.alpha) < 1e-9 and abs(g.mll - o.mll) < 1e-11 * abs(o.mll)
        if eng in res:
            assert np.array_equal(res[eng][0], U) and np.array_equal(res[eng][1], g.alpha)      # bitwise reproducible
        res[eng] = (U, g.alpha.copy())
    assert relmax(res[1][0], res[0][0]) < 1e-12 and relmax(res[2][0], res[0][0]) < 1e-12


@pytest.mark.parametrize("kern,D,N,M", [("Mat52Ard", 6, 2048, 3000), ("SEArd", 32, 1500, 700), ("Mat12Iso", 3, 130, 129), ("SEArd", 2, 700, 4500)])
def test_acquisition_engines_agree(bo, kern, D, N, M):
    """K6 on tcgen05 (int8-slice GEMM against W = L^-1, the default) and on the FP64 tensor pipe (blocked DMMA solves) against the
    oracle and against each other; candidates include training points, where sigma^2 = k** - v'v cancels most."""
    rng, o, g, X, y = make_pair(bo, kern, "MeanConst", D, N, seed=31 * N + D)
    Xs = rng.random((D, M)); Xs[:, :40] = X[:, :40]; Xs[:, 40:60] = X[:, 40:60] + 1e-8
    mo, vo = o.predict(Xs)
    tau = float(np.quantile(y, 0.9))
    a, gr = orc.acq_grad(o, "EI", (tau,), Xs)
    out = {}
    for eng in (1, 0):
        g.set_acq_engine(eng)
        r = g.acquire("EI", (tau,), Xs, want_grad=True, want_mu_var=True)
        assert close(r["mu"], mo, RTOL_POST) and close(r["var"], vo, RTOL_POST, 1e-13) and close(r["values"], a, RTOL_ACQ), eng
        assert r["best_index"] == orc.first_strict_argmax_np(a), eng
        assert relmax(r["grad"], gr) < 1e-8, eng
        r2 = g.acquire("EI", (tau,), Xs)                                    # value-only launch: same bits, same selection
        assert np.array_equal(r2["values"], r["values"]) and r2["best_index"] == r["best_index"], eng
        out[eng] = r
    assert np.max(np.abs(out[1]["var"] - out[0]["var"])) < 1e-12 * o.sf2 and relmax(out[1]["mu"], out[0]["mu"]) < 1e-12
    assert relmax(out[1]["grad"], out[0]["grad"]) < 1e-9


@pytest.mark.parametrize("N", [1920, 2560, 4096])
def test_factor_many_panels(bo, N):
    """blocked Cholesky with many 128-panels (persistent trailing update walks several tiles per CTA, look-ahead streams)."""
    rng, o, g, X, y = make_pair(bo, "SEArd", "MeanConst", 8, N, seed=N)
    assert g.jitter_tries == 0
    assert relmax(g.factor, o.U) < 1e-11
    assert relmax(g.alpha, o.alpha) < 1e-9 and abs(g.mll - o.mll) < 1e-11 * abs(o.mll)
    g.fit(X, y)                                   # refit: bitwise reproducible
    assert np.array_equal(g.factor, g.factor) and relmax(g.factor, o.U) < 1e-11


def test_batched_equals_per_point_exactly(bo):
    """reference test/acquisitionfunctions.jl:1-12: acfunc(X)[1] == acfunc(X[:,1]) for PI/EI/UCB/MI; length 2 for all."""
    rng = np.random.default_rng(0)
    g = bo.B200GPE.from_data(rng.random((3, 4)), rng.random(4), bo.MeanZero(), bo.SEIso(0.0, 0.0))
    x = rng.random((3, 2))
    for ac in [bo.ProbabilityOfImprovement(), bo.ExpectedImprovement(), bo.UpperConfidenceBound(), bo.ThompsonSamplingSimple(),
               bo.MutualInformation()]:
        f = bo.acquisitionfunction(ac, g)
        v = f(x)
        assert len(v) == 2
        if not isinstance(ac, bo.ThompsonSamplingSimple):
            assert v[0] == f(x[:, 0])
    # position independence inside and across tiles, and across batch sizes
    rng, o, g, X, y = make_pair(bo, "Mat52Ard", "MeanConst", 5, 300, 1)
    Xs = rng.random((5, 200))
    full = g.acquire("EI", (0.2,), Xs, want_grad=True)
    for j in (0, 1, 63, 64, 130, 199):
        one = g.acquire("EI", (0.2,), Xs[:, j], want_grad=True)
        assert one["values"][0] == full["values"][j] and np.array_equal(one["grad"][:, 0], full["grad"][:, j])
    perm = rng.permutation(200)
    assert np.array_equal(g.acquire("EI", (0.2,), Xs[:, perm])["values"], full["values"][perm])


def test_sharding_invariance_and_tie_break(bo):
    """shards of one candidate matrix agree with the whole (values, TS noise, global arg-max); ties -> lowest index."""
    from b200bo.dist import shard_bounds, select_best
    rng, o, g, X, y = make_pair(bo, "SEArd", "MeanConst", 4, 200, 2)
    Xs = rng.random((4, 1000))
    Xs[:, 700] = Xs[:, 123]; Xs[:, 5] = Xs[:, 123]                   # exact duplicates -> exact ties
    for kind, par in [("UCB", (2.0,)), ("TS", ())]:
        whole = g.acquire(kind, par, Xs, seed=9)
        parts = []
        for r in range(3):
            lo, hi = shard_bounds(1000, 3, r)
            parts.append(g.acquire(kind, par, Xs[:, lo:hi], seed=9, idx_offset=lo))
        assert np.array_equal(np.concatenate([p["values"] for p in parts]), whole["values"])
        assert select_best([p["best_value"] for p in parts], [p["best_index"] for p in parts]) == (whole["best_value"], whole["best_index"])
    flat = np.zeros((4, 300)) + 0.5                                    # all candidates identical: first index must win
    assert g.acquire("EI", (0.0,), flat)["best_index"] == 0
    assert g.acquire("EI", (0.0,), flat, idx_offset=77)["best_index"] == 77


def test_edge_cases(bo):
    g = bo.B200GPE(2, mean=bo.MeanConst(-1.5), kernel=bo.SEArd([0.0, 0.0], 0.5), logNoise=-2.0, capacity=10)
    # empty model = prior (setparams!/acquire on a model without data, quirks 4-5)
    assert bo.dims(g) == (2, 0) and bo.maxy(g) == -np.inf
    mu, var = g.predict(np.array([[0.1, 0.2], [0.3, 0.4]]))
    assert np.all(mu == -1.5) and np.allclose(var, np.exp(1.0), rtol=1e-15)
    r = g.acquire("EI", (-np.inf,), np.zeros((2, 3)))                  # tau = -Inf: value = +Inf everywhere, first wins
    assert r["best_index"] == 0 and r["best_value"] == np.inf
    r = g.acquire("UCB", (1.0,), np.zeros((2, 0)))                     # M = 0: nothing wins
    assert r["best_index"] == -1 and r["best_value"] == -np.inf
    # elastic growth beyond capacity, append == fit
    rng = np.random.default_rng(4)
    X = rng.random((2, 300)); y = rng.standard_normal(300)
    g.fit(X[:, :7], y[:7])
    bo.update(g, X[:, 7:140], y[7:140])
    bo.update(g, X[:, 140:], y[140:])
    assert bo.dims(g) == (2, 300) and np.array_equal(g.x, X) and np.array_equal(g.y, y) and bo.maxy(g) == y.max()
    g2 = bo.B200GPE(2, mean=bo.MeanConst(-1.5), kernel=bo.SEArd([0.0, 0.0], 0.5), logNoise=-2.0, capacity=300)
    g2.fit(X, y)
    assert np.array_equal(g.alpha, g2.alpha) and g.mll == g2.mll
    bo.update(g, np.zeros((2, 0)), np.zeros(0))                        # empty y: refit on current data (gp.jl:13-14)
    assert g.mll == g2.mll
    r = g.acquire("MaxMean", (), np.full((2, 4), np.nan))              # NaN never wins (acquisition.jl:62 `f > maxf`)
    assert r["best_index"] == -1 and r["best_value"] == -np.inf
    nan1 = rng.random((2, 5)); nan1[0, 1] = np.nan
    r = g.acquire("UCB", (1.0,), nan1)
    assert r["best_index"] in (0, 2, 3, 4) and np.isnan(r["values"][1]) and r["best_index"] == int(np.nanargmax(r["values"]))
    # argument errors surface as errors, not crashes
    with pytest.raises(ValueError):
        g.predict(np.zeros((3, 2)))
    with pytest.raises(bo._lib.B200BOError):
        g.acquire("TS", (), np.zeros((2, 2)), want_grad=True)
    with pytest.raises(bo._lib.B200BOError):
        g.acquire("MI", (1.0,), np.zeros((2, 2)))


def test_clamp_and_zero_variance_branches(bo):
    """sigma^2 clamps at 0 on (near-)noise-free training points; the exact-zero branches of EI/PI apply (quirk 2)."""
    X = np.array([[0.0, 0.5, 1.0]]); y = np.array([0.0, 1.0, 0.5])
    o = orc.GPOracle(1, "SEIso", "MeanZero", ll=[-1.0], lsigma=0.0, lognoise=-20.0).fit(X, y)
    g = bo.B200GPE(1, mean=bo.MeanZero(), kernel=bo.SEIso(-1.0, 0.0), logNoise=-20.0, capacity=3)
    g.fit(X, y)
    Xs = np.array([[0.0, 0.5, 1.0, 0.25]])
    mu, var = g.predict(Xs)
    mo, vo = o.predict(Xs)
    assert np.allclose(mu, mo, rtol=1e-6, atol=1e-9) and np.all(var >= 0.0) and np.all(var[:3] < 1e-8)
    for kind in ("EI", "PI"):
        r = g.acquire(kind, (0.7,), Xs, want_grad=True)
        a = orc.acq_value(kind, (0.7,), mu, var)                       # functor applied to the device's own (mu, var)
        assert np.allclose(r["values"], a, rtol=1e-5, atol=1e-12) and np.all(np.isfinite(r["grad"]))


def test_jitter_retry_matches_make_posdef(bo):
    X = np.linspace(0.0, 1.0, 60)[None, :]; y = np.sin(4 * X[0])
    o = orc.GPOracle(1, "SEIso", "MeanZero", ll=[3.0], lsigma=0.0, lognoise=-30.0).fit(X, y)
    g = bo.B200GPE(1, mean=bo.MeanZero(), kernel=bo.SEIso(3.0, 0.0), logNoise=-30.0, capacity=60)
    g.fit(X, y)
    assert g.jitter_tries == o.jitter_tries == 1
    assert abs(g.mll - o.mll) < 1e-6 * abs(o.mll)


def test_full_size_properties_config2(bo):
    """BASELINE config 2 shape (Hartmann-6, N=2048, Mat52Ard, UCB/Brochu, M=65536): size-independent properties
    plus the oracle on a random subsample of the candidates."""
    rng = np.random.default_rng(2)
    D, N, M = 6, 2048, 65536
    X = rng.random((D, N)); y = -orc.hartmann6(X)
    g = bo.B200GPE(D, mean=bo.MeanConst(0.0), kernel=bo.Mat52Ard(np.zeros(D), 0.0), logNoise=-2.0, capacity=N)
    g.fit(X, y)
    o = orc.GPOracle(D, "Mat52Ard", "MeanConst", ll=np.zeros(D), lsigma=0.0, lognoise=-2.0, beta=0.0).fit(X, y)
    assert relmax(g.alpha, o.alpha) < 1e-9 and abs(g.mll - o.mll) < 1e-11 * abs(o.mll)
    # (1) mu(x_i) = y_i - noise * alpha_i at the training points (K alpha = (y - m) - noise alpha)
    mu_tr, var_tr = g.predict(X)
    noise = np.exp(-4.0) + orc.EPS
    assert np.allclose(mu_tr, y - noise * g.alpha, rtol=0, atol=1e-10)
    assert np.all(var_tr >= 0) and np.all(var_tr <= 1.0)
    # (2) the whole batch: arg-max consistent with its own values; subsample against the oracle
    Xs = orc.latin_hypercube_sampling(np.zeros(D), np.ones(D), M, np.random.default_rng(20))
    beta = orc.brochu_beta(D, N)
    r = g.acquire("UCB", (beta,), Xs, want_mu_var=True)
    assert r["best_index"] == orc.first_strict_argmax_np(r["values"]) and r["best_value"] == r["values"][r["best_index"]]
    sub = np.sort(np.random.default_rng(21).choice(M, 384, replace=False))
    sub[0] = r["best_index"]
    mo, vo = o.predict(Xs[:, sub])
    assert close(r["mu"][sub], mo, RTOL_POST) and close(r["var"][sub], vo, RTOL_POST)
    assert close(r["values"][sub], orc.acq_value("UCB", (beta,), mo, vo), RTOL_ACQ)
    # (3) sharding over 8 ranks reproduces the whole
    from b200bo.dist import shard_bounds, select_best
    parts = [g.acquire("UCB", (beta,), Xs[:, lo:hi], idx_offset=lo, want_values=False) for lo, hi in (shard_bounds(M, 8, k) for k in range(8))]
    assert select_best([p["best_value"] for p in parts], [p["best_index"] for p in parts]) == (r["best_value"], r["best_index"])
    assert g.launch_count > 0


def _bumps(X, rng):
    c = rng.random((X.shape[0], 8))
    return sum(np.exp(-0.5 * np.sum((X - c[:, k:k + 1]) ** 2, axis=0) / 0.15) for k in range(8)) + np.exp(-2.0) * rng.standard_normal(X.shape[1])


def test_full_size_properties_config3(bo):
    """BASELINE configs[2] at its full model size (D=32, N=4096, EI + gradient) on a slice of the candidate sweep:
    size-independent properties + an oracle subsample + a finite-difference check of the fused gradient."""
    rng = np.random.default_rng(3)
    D, N, M = 32, 4096, 8192
    X = rng.random((D, N)); y = _bumps(X, rng)
    ll = np.full(D, np.log(np.sqrt(D) * 0.25))
    g = bo.B200GPE(D, mean=bo.MeanConst(0.0), kernel=bo.SEArd(ll, 0.0), logNoise=-2.0, capacity=N)
    g.fit(X, y)
    assert g.jitter_tries == 0
    o = orc.GPOracle(D, "SEArd", "MeanConst", ll=ll, lsigma=0.0, lognoise=-2.0, beta=0.0).fit(X, y)
    assert relmax(g.alpha, o.alpha) < 1e-8 and abs(g.mll - o.mll) < 1e-10 * abs(o.mll)
    noise = np.exp(-4.0) + orc.EPS
    mu_tr, var_tr = g.predict(X[:, :1024])
    assert np.allclose(mu_tr, y[:1024] - noise * g.alpha[:1024], rtol=0, atol=1e-9) and np.all(var_tr >= 0)
    Xs = orc.latin_hypercube_sampling(np.zeros(D), np.ones(D), M, np.random.default_rng(30))
    tau = float(y.max())
    r = g.acquire("EI", (tau,), Xs, want_grad=True, want_mu_var=True)
    assert r["best_index"] == orc.first_strict_argmax_np(r["values"])
    sub = np.sort(np.random.default_rng(31).choice(M, 96, replace=False)); sub[0] = r["best_index"]
    a, gr = orc.acq_grad(o, "EI", (tau,), Xs[:, sub])
    mo, vo = o.predict(Xs[:, sub])
    assert close(r["mu"][sub], mo, RTOL_POST) and close(r["var"][sub], vo, RTOL_POST) and close(r["values"][sub], a, RTOL_ACQ)
    assert relmax(r["grad"][:, sub], gr) < 1e-7
    # central finite differences of the device value along random directions agree with the device gradient
    k = int(sub[1]); u = np.random.default_rng(32).standard_normal(D); u /= np.linalg.norm(u); h = 1e-5
    pm = g.acquire("EI", (tau,), np.stack([Xs[:, k] + h * u, Xs[:, k] - h * u], axis=1))["values"]
    fd = (pm[0] - pm[1]) / (2 * h)
    assert abs(fd - r["grad"][:, k] @ u) <= 1e-5 * max(abs(fd), np.abs(r["grad"][:, k]).max())
    # the gradient launch and the value-only launch give the same values and the same selection
    r0 = g.acquire("EI", (tau,), Xs)
    assert np.array_equal(r0["values"], r["values"]) and r0["best_index"] == r["best_index"]


def test_full_size_properties_config5_shard(bo):
    """BASELINE configs[4] at its full model size (D=16, N=8192, ThompsonSamplingSimple) on a slice of one rank's shard: the
    Philox stream is keyed by the GLOBAL candidate index, so shards of different shapes agree bit for bit."""
    rng = np.random.default_rng(5)
    D, N, M = 16, 8192, 4096
    X = rng.random((D, N)); y = _bumps(X, rng)
    ll = np.full(D, np.log(np.sqrt(D) * 0.25))
    g = bo.B200GPE(D, mean=bo.MeanConst(0.0), kernel=bo.SEArd(ll, 0.0), logNoise=-2.0, capacity=N)
    g.fit(X, y)
    assert g.jitter_tries == 0
    noise = np.exp(-4.0) + orc.EPS
    mu_tr, var_tr = g.predict(X[:, -512:])
    assert np.allclose(mu_tr, y[-512:] - noise * g.alpha[-512:], rtol=0, atol=1e-9) and np.all(var_tr >= 0) and np.all(var_tr <= 1.0)
    Xs = orc.latin_hypercube_sampling(np.zeros(D), np.ones(D), M, np.random.default_rng(51))
    off = 3 * 131072                                            # rank 3 of 8
    r = g.acquire("TS", (), Xs, seed=50, idx_offset=off, want_mu_var=True)
    eps = orc.philox_normal(50, off + np.arange(M))
    assert close(r["values"], orc.acq_value("TS", (), r["mu"], r["var"], eps=eps), RTOL_ACQ)
    assert r["best_index"] - off == orc.first_strict_argmax_np(r["values"])
    halves = [g.acquire("TS", (), Xs[:, lo:hi], seed=50, idx_offset=off + lo) for lo, hi in ((0, 1000), (1000, M))]
    assert np.array_equal(np.concatenate([h_["values"] for h_ in halves]), r["values"])
    # the oracle at the full model size on a subsample of the shard (dpotrf at N = 8192 takes a few seconds on the host)
    o = orc.GPOracle(D, "SEArd", "MeanConst", ll=ll, lsigma=0.0, lognoise=-2.0, beta=0.0).fit(X, y)
    assert relmax(g.alpha, o.alpha) < 1e-8 and abs(g.mll - o.mll) < 1e-10 * abs(o.mll)
    sub = np.sort(np.random.default_rng(52).choice(M, 64, replace=False)); sub[0] = r["best_index"] - off
    mo, vo = o.predict(Xs[:, sub])
    assert close(r["mu"][sub], mo, RTOL_POST) and close(r["var"][sub], vo, RTOL_POST)
    assert close(r["values"][sub], orc.acq_value("TS", (), mo, vo, eps=eps[sub]), RTOL_ACQ)


def test_full_size_config4_map_sweep(bo):
    """BASELINE configs[3] at full size: N=4096, D=8, the 8 x 8 grid of (logNoise, common SEArd length-scale) settings through ONE
    b200bo_mll_sweep call with gradients; the diagonal of the grid (8 settings, every logNoise and every length-scale once) against
    the oracle's mll / dmll (closure of optimizemodel!, reference src/models/gp.jl:59-64)."""
    rng = np.random.default_rng(4)
    D, N = 8, 4096
    X = rng.random((D, N)); y = _bumps(X, rng)
    g = bo.B200GPE(D, mean=bo.MeanConst(0.0), kernel=bo.SEArd(np.zeros(D), 0.0), logNoise=-2.0, capacity=N)
    g.fit(X, y)
    th0 = g.get_params()
    grid = [(ln, l) for ln in np.linspace(-3.0, 0.0, 8) for l in np.linspace(-1.5, 0.5, 8)]
    Theta = np.stack([np.concatenate([[ln, 0.0], np.full(D, l), [0.0]]) for ln, l in grid], axis=1)     # [logNoise, beta, ll.., lsigma]
    mll, dmll = g.mll_sweep(Theta)
    assert mll.shape == (64,) and dmll.shape == (D + 3, 64) and np.all(np.isfinite(mll)) and np.all(np.isfinite(dmll))
    o = orc.GPOracle(D, "SEArd", "MeanConst", ll=np.zeros(D), lsigma=0.0, lognoise=-2.0, beta=0.0).fit(X, y)
    for k in range(0, 64, 9):
        mo, go = o.mll_dmll(Theta[:, k])
        assert abs(mll[k] - mo) <= 1e-9 * abs(mo), k
        assert relmax(dmll[:, k], go) < 1e-7, k
    assert np.array_equal(g.get_params(), th0)                                     # the sweep leaves the model where it was
    # a sweep of a subset returns the same numbers as the same settings inside the full sweep
    m2, d2 = g.mll_sweep(Theta[:, [9, 54]])
    assert np.array_equal(m2, mll[[9, 54]]) and np.array_equal(d2, dmll[:, [9, 54]])


@pytest.mark.parametrize("kern", ["SEArd", "Mat52Ard", "Mat12Ard"])
def test_kmat_uncentred_inputs_small_lengthscale(bo, kern):
    """K1 takes its distances in Gram form (z_i.z_j - |z_i|^2/2 - |z_j|^2/2), whose rounding error grows with |z|^2.  Branin's box
    [-5,10] x [0,15] (reference test/branin.jl:1-5) with l = e^-2 puts |z|^2 at ~1.7e4 uncentred: check K element by element (relative, wherever the
    entry is not negligible) and the posterior against the difference-form oracle."""
    rng = np.random.default_rng(17)
    D, N, M = 2, 900, 400
    lb, ub = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    X = lb[:, None] + (ub - lb)[:, None] * rng.random((D, N))
    y = -orc.branin(X[0], X[1]) / 50.0
    ll = np.full(D, -2.0)
    g = bo.B200GPE(D, mean=bo.MeanConst(-1.0), kernel=bo.gp._Kernel(kern, ll, 0.0), logNoise=-2.0, capacity=N)
    g.fit(X, y)
    o = orc.GPOracle(D, kern, "MeanConst", ll=ll, lsigma=0.0, lognoise=-2.0, beta=-1.0).fit(X, y)
    K = g.kmat()
    S = o.cov(o.X, o.X); S[np.diag_indices(N)] += np.exp(-4.0) + orc.EPS
    big = S > 1e-8
    assert float(np.max(np.abs(K - S)[big] / S[big])) < 1e-11 and float(np.max(np.abs(K - S)[~big])) < 1e-18
    assert relmax(g.alpha, o.alpha) < 1e-9
    Xs = lb[:, None] + (ub - lb)[:, None] * rng.random((D, M))
    mu, var = g.predict(Xs)
    mo, vo = o.predict(Xs)
    assert close(mu, mo, 1e-9, 1e-12) and close(var, vo, 1e-9, 1e-12)


def test_duplicate_points_mat12(bo):
    """`repetitions > 1` appends hcat(fill(x, rep)) (reference src/BayesianOptimization.jl:194-196) and the search re-proposes points:
    exact duplicates must give r = 0 exactly in K1 (Mat12: k = exp(-sqrt(r2)) turns a rounding residual of 1e-16 |z|^2 into 1e-8),
    consistently between a full refit, the elastic append and the oracle."""
    rng = np.random.default_rng(23)
    D, N0 = 3, 300
    X0 = 4.0 + 3.0 * rng.random((D, N0)); y0 = np.sin(X0.sum(0))
    dup = np.repeat(X0[:, 5:9], 2, axis=1)                                   # every point twice, and each already in the model
    X = np.hstack([X0, dup]); y = np.concatenate([y0, np.sin(dup.sum(0)) + 0.01 * rng.standard_normal(dup.shape[1])])
    ll = np.full(D, 0.3)
    mk = lambda: bo.B200GPE(D, mean=bo.MeanZero(), kernel=bo.gp._Kernel("Mat12Ard", ll, 0.0), logNoise=-2.0, capacity=N0 + 128)
    ref = mk(); ref.fit(X, y)
    o = orc.GPOracle(D, "Mat12Ard", "MeanZero", ll=ll, lsigma=0.0, lognoise=-2.0).fit(X, y)
    K = ref.kmat()
    S = o.cov(o.X, o.X); S[np.diag_indices(X.shape[1])] += np.exp(-4.0) + orc.EPS
    assert float(np.max(np.abs(K - S) / S)) < 1e-12
    assert K[5, N0] == 1.0 and K[5, N0 + 1] == 1.0 and K[N0, N0 + 1] == 1.0   # sf2 * exp(-0) exactly
    g = mk(); g.fit(X0, y0)
    bo.update(g, X[:, N0:], y[N0:])                                          # elastic append of the duplicates
    assert relmax(g.factor, ref.factor) < 1e-11 and relmax(g.factor, o.U) < 1e-11
    assert relmax(g.alpha, o.alpha) < 1e-9 and abs(g.mll - o.mll) < 1e-10 * abs(o.mll)


def test_linearity_in_y_large(bo):
    """posterior mean is linear in y (MeanZero): mu[y1 + y2] = mu[y1] + mu[y2]; variance does not depend on y."""
    rng = np.random.default_rng(8)
    D, N, M = 8, 1536, 4096
    X = rng.random((D, N)); y1 = rng.standard_normal(N); y2 = np.cos(X.sum(0))
    Xs = rng.random((D, M))
    out = []
    for yy in (y1, y2, y1 + y2):
        g = bo.B200GPE(D, mean=bo.MeanZero(), kernel=bo.SEArd(np.full(D, np.log(0.7)), 0.0), logNoise=-2.0, capacity=N)
        g.fit(X, yy)
        out.append(g.predict(Xs))
    assert np.allclose(out[0][0] + out[1][0], out[2][0], rtol=0, atol=1e-9 * np.abs(out[2][0]).max())
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][1], out[2][1])


def test_device_lhs_matches_restatement_and_is_stratified(bo):
    """src/utils.jl:101-120 on device: every stratum used once per dimension; bit-exact vs the oracle restatement of the
    keyed permutation; shards of one design agree with the whole."""
    g = bo.B200GPE(3, mean=bo.MeanZero(), kernel=bo.SEArd(np.zeros(3), 0.0), capacity=8)
    lb, ub = np.array([-5.0, 0.0, 2.0]), np.array([10.0, 15.0, 2.0])
    for n, seed in [(1, 0), (7, 3), (64, 1), (1000, 123456789012345), (4097, 7)]:
        X = g.lhs(lb, ub, n, seed=seed)
        assert X.shape == (3, n) and np.array_equal(X, orc.lhs_device(lb, ub, n, seed))
        for d in range(2):
            strata = np.floor((X[d] - lb[d]) / ((ub[d] - lb[d]) / n)).astype(int)
            assert sorted(strata) == list(range(n))
        assert np.all(X[2] == 2.0)
    whole = g.lhs(lb, ub, 1000, seed=5)
    parts = [g.lhs(lb, ub, 1000, seed=5, offset=o, n_local=m) for o, m in [(0, 333), (333, 334), (667, 333)]]
    assert np.array_equal(np.hstack(parts), whole)
    with pytest.raises(bo._lib.B200BOError):
        g.lhs(ub + 1, lb, 10)


def test_acquire_lhs_equals_host_sweep(bo):
    rng, o, g, X, y = make_pair(bo, "Mat52Ard", "MeanConst", 4, 300, 31)
    lb, ub = np.zeros(4), np.ones(4)
    for kind, par in [("EI", (float(np.quantile(y, 0.9)),)), ("TS", ())]:
        r = g.acquire_lhs(kind, par, lb, ub, 5000, lhs_seed=11, ts_seed=12, want_values=True)
        Xs = g.lhs(lb, ub, 5000, seed=11)
        r0 = g.acquire(kind, par, Xs, seed=12)
        assert np.array_equal(r["values"], r0["values"]) and r["best_index"] == r0["best_index"] and np.array_equal(r["best_x"], r0["best_x"])
        a = g.acquire_lhs(kind, par, lb, ub, 5000, lhs_seed=11, ts_seed=12, offset=0, n_local=2500)
        b = g.acquire_lhs(kind, par, lb, ub, 5000, lhs_seed=11, ts_seed=12, offset=2500, n_local=2500)
        from b200bo.dist import select_best
        assert select_best([a["best_value"], b["best_value"]], [a["best_index"], b["best_index"]]) == (r["best_value"], r["best_index"])


def test_batched_ascent_against_restatement(bo):
    rng, o, g, X, y = make_pair(bo, "SEArd", "MeanConst", 3, 120, 41)
    lb, ub = np.zeros(3), np.ones(3)
    X0 = rng.random((3, 100))
    for kind, par in [("UCB", (2.0,)), ("EI", (float(np.quantile(y, 0.8)),)), ("MaxMean", ())]:
        r = g.acquire_ascent(kind, par, X0, lb, ub, steps=15, step0=0.05)
        Xo, Fo = orc.ascent(o, kind, par, X0, lb, ub, steps=15, step0=0.05)
        f0 = g.acquire(kind, par, X0)["values"]
        assert np.all(r["values"] >= f0) and np.all(r["X"] >= lb[:, None]) and np.all(r["X"] <= ub[:, None])
        assert np.mean(r["values"] > f0 + 1e-9 * np.abs(f0)) > 0.8                      # ascent actually climbs
        agree = np.abs(r["values"] - Fo) <= 1e-6 * np.abs(Fo) + 1e-10                    # same trajectories up to accept/reject ties
        assert agree.mean() > 0.97
        chk = g.acquire(kind, par, r["X"])["values"]                                     # returned points reproduce the returned values
        assert close(chk, r["values"], 1e-9)
        assert r["best_index"] == orc.first_strict_argmax_np(r["values"]) and r["best_value"] == r["values"][r["best_index"]]
    with pytest.raises(bo._lib.B200BOError):
        g.acquire_ascent("TS", (), X0, lb, ub)


@pytest.mark.parametrize("kern,N0", [("SEArd", 126), ("Mat52Ard", 250), ("Mat32Iso", 128), ("SEArd", 1000)])
def test_elastic_append_matches_refit(bo, kern, N0):
    """update!(model::ElasticGPE, x, y) = append! (gp.jl:11): rank-1 factor extension on the device, across a 128-block
    boundary, against a full refactor and the oracle."""
    D = 4
    rng = np.random.default_rng(N0)
    X = rng.random((D, N0 + 9)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N0 + 9)
    ll = np.full(1 if kern.endswith("Iso") else D, np.log(0.5))
    mk = lambda: bo.B200GPE(D, mean=bo.MeanConst(0.2), kernel=bo.gp._Kernel(kern, ll, 0.1), logNoise=-2.0, capacity=N0 + 200)
    g = mk(); g.fit(X[:, :N0], y[:N0])
    l0 = g.launch_count
    for j in range(N0, N0 + 5):
        bo.update(g, X[:, j], y[j:j + 1])                     # one point per BO iteration
    bo.update(g, X[:, N0 + 5:], y[N0 + 5:])                    # a small batch
    assert bo.dims(g) == (D, N0 + 9) and np.array_equal(g.y, y)
    ref = mk(); ref.fit(X, y)
    o = orc.GPOracle(D, kern, "MeanConst", ll=ll, lsigma=0.1, lognoise=-2.0, beta=0.2).fit(X, y)
    assert relmax(g.factor, o.U) < 1e-11 and relmax(g.alpha, o.alpha) < 1e-9 and abs(g.mll - o.mll) < 1e-10 * abs(o.mll)
    assert relmax(g.factor, ref.factor) < 1e-12
    Xs = rng.random((D, 300))
    r, r0 = g.acquire("EI", (0.3,), Xs, want_grad=True, want_mu_var=True), ref.acquire("EI", (0.3,), Xs, want_grad=True, want_mu_var=True)
    assert close(r["mu"], r0["mu"], 1e-9) and close(r["var"], r0["var"], 1e-8) and r["best_index"] == r0["best_index"]
    assert relmax(r["grad"], r0["grad"]) < 1e-8
    g.set_params(g.get_params() + 0.1)                         # hyper-parameters change: the next update refactors
    bo.update(g, X[:, :1], y[:1])
    ref.set_params(ref.get_params() + 0.1); ref.fit(np.hstack([X, X[:, :1]]), np.concatenate([y, y[:1]]))
    assert relmax(g.alpha, ref.alpha) < 1e-10


def test_kmat_dev_unaligned_target_uses_the_16_byte_store_path(bo):
    """b200bo_kmat_dev into a caller buffer that is only 16-byte aligned with ld = 2 (mod 4): K1's 128-bit store instantiation
    (the 32-byte STG.E.256 path needs 32-byte alignment and ld % 4 == 0) must give the same matrix, ragged N included."""
    import ctypes as C
    import torch
    from b200bo import _lib
    rng = np.random.default_rng(9)
    D, N = 5, 333
    X = rng.random((D, N)); y = rng.standard_normal(N)
    g = bo.B200GPE(D, mean=bo.MeanZero(), kernel=bo.Mat32Ard(np.full(D, -0.3), 0.2), logNoise=-1.5, capacity=N)
    g.fit(X, y)
    K = g.kmat()
    for ld, off in ((336, 0), (334, 2)):                                        # 32-byte aligned, ld % 4 == 0  /  16-byte aligned, ld % 4 == 2
        buf = torch.zeros(ld * N + 8, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()                                                # the fill runs on torch's stream, K1 on the library's
        ptr = buf.data_ptr() + 8 * off
        assert ptr % 32 == (16 if off else 0)
        _lib.check(_lib.lib.b200bo_kmat_dev(g._h, C.c_void_p(ptr), ld), g._h)
        _lib.check(_lib.lib.b200bo_sync(g._h), g._h)
        out = buf[off:off + ld * N].view(N, ld)[:, :N].cpu().numpy()
        assert np.array_equal(out, K), (ld, off)


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
def test_multi_gpu_handle_equals_single_gpu_bit_for_bit(bo, n_gpus):
    """b200bo_create_multi: the candidate columns shard over n_gpus replicas INSIDE the library (contiguous blocks, one 272 B/rank
    ncclAllGather, deterministic merge on the device); values, posterior, gradient, selected index and point equal the single-GPU call
    bit for bit; the settings of a MAP sweep shard the same way (reference loop: src/acquisition.jl:58-66)."""
    if _ngpu() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    rng = np.random.default_rng(77)
    D, N, M = 5, 700, 5000
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    Xs = rng.random((D, M)); Xs[:, 4321] = Xs[:, 17]                                   # an exact tie across shards: lowest index wins
    mk = lambda **kw: bo.B200GPE(D, mean=bo.MeanConst(0.1), kernel=bo.Mat52Ard(np.full(D, -0.5), 0.2), logNoise=-2.0, capacity=N + 50, **kw)
    g1 = mk(); g1.fit(X, y)
    gm = mk(n_gpus=n_gpus); gm.fit(X, y)
    assert gm.num_gpus == n_gpus
    tau = float(np.quantile(y, 0.9))
    for kind, par, kw in (("EI", (tau,), dict(want_grad=True, want_mu_var=True)), ("UCB", (2.0,), {}), ("TS", (), dict(seed=9, idx_offset=1000))):
        a, b = g1.acquire(kind, par, Xs, **kw), gm.acquire(kind, par, Xs, **kw)
        assert np.array_equal(a["values"], b["values"]) and a["best_index"] == b["best_index"] and a["best_value"] == b["best_value"], kind
        assert np.array_equal(a["best_x"], b["best_x"]), kind
        if "want_grad" in kw:
            assert np.array_equal(a["grad"], b["grad"]) and np.array_equal(a["mu"], b["mu"]) and np.array_equal(a["var"], b["var"])
    m1, v1 = g1.predict(Xs[:, :999]); m2, v2 = gm.predict(Xs[:, :999])
    assert np.array_equal(m1, m2) and np.array_equal(v1, v2)
    bo.update(g1, Xs[:, :3], np.zeros(3)); bo.update(gm, Xs[:, :3], np.zeros(3))       # elastic append on every replica
    assert np.array_equal(g1.acquire("EI", (tau,), Xs)["values"], gm.acquire("EI", (tau,), Xs)["values"])
    th = g1.get_params()
    Theta = np.stack([th, th + 0.1, th - 0.2, th + 0.05, th - 0.1], axis=1)
    ma, da = g1.mll_sweep(Theta); mb, db = gm.mll_sweep(Theta)
    assert np.array_equal(ma, mb) and np.array_equal(da, db)
    lb, ub = np.zeros(D), np.ones(D)
    ra, rb = g1.acquire_lhs("EI", (tau,), lb, ub, 6000, lhs_seed=3), gm.acquire_lhs("EI", (tau,), lb, ub, 6000, lhs_seed=3)
    assert ra["best_index"] == rb["best_index"] and ra["best_value"] == rb["best_value"] and np.array_equal(ra["best_x"], rb["best_x"])


def test_cusolver_cublas_cross_check(bo):
    """A second, independent GPU-side oracle (tests only): cuSOLVER dpotrf + cuBLAS dtrsm through torch.linalg in FP64 on
    the library's own Sigma.  Factor, alpha, log-determinant and the posterior of the hand-written path against the vendor libraries."""
    import torch
    rng, o, g, X, y = make_pair(bo, "Mat52Ard", "MeanConst", 6, 1500, seed=4242)
    S = torch.from_numpy(g.kmat()).cuda()
    L = torch.linalg.cholesky(S)                                               # cuSOLVER potrf
    U = L.T.cpu().numpy()
    assert relmax(g.factor, U) < 1e-11
    r = torch.from_numpy(y - 0.3).cuda()
    alpha = torch.cholesky_solve(r[:, None], L)[:, 0]                          # cuBLAS trsm x 2
    assert relmax(g.alpha, alpha.cpu().numpy()) < 1e-9
    mll = -0.5 * (float(r @ alpha) + 2.0 * float(torch.log(torch.diagonal(L)).sum()) + y.size * np.log(2 * np.pi))
    assert abs(g.mll - mll) <= 1e-11 * abs(mll)
    Xs = rng.random((6, 2000)); Xs[:, :20] = X[:, :20]
    Ks = torch.from_numpy(o.cov(o.X, Xs)).cuda()
    V = torch.linalg.solve_triangular(L, Ks, upper=False)                      # cuBLAS trsm
    var = np.maximum(o.sf2 - (V * V).sum(0).cpu().numpy(), 0.0)
    mu = 0.3 + (Ks.T @ alpha).cpu().numpy()
    for eng in (1, 0):
        g.set_acq_engine(eng)
        m, v = g.predict(Xs)
        assert close(m, mu, 1e-9, 1e-12) and close(v, var, 1e-8, 1e-13), eng


def test_device_sobol_matches_joe_kuo_and_the_reference_iterator(bo):
    """b200bo_sobol (reference ScaledSobolIterator, src/utils.jl:64-87): bit-equal to the unscrambled Joe-Kuo sequence SciPy
    and Sobol.jl ship, any block of indices; the reference iterator's points are the block starting at 1 + 2^floor(log2(N+1))."""
    from scipy.stats import qmc
    for D in (2, 8, 32):
        g = bo.B200GPE(D, capacity=128)
        lb = -np.arange(1.0, D + 1.0); ub = 2.0 + np.arange(D)
        pts = g.sobol(lb, ub, 9, 700)
        s = qmc.Sobol(D, scramble=False, bits=32); s.fast_forward(9)
        u = s.random(700)
        assert np.array_equal(pts, (lb + (ub - lb) * u).T), D
    g = bo.B200GPE(2, capacity=128)
    host = np.array(list(bo.ScaledSobolIterator([-5.0, 0.0], [10.0, 15.0], 10))).T
    assert np.array_equal(g.sobol([-5.0, 0.0], [10.0, 15.0], 1 + 8, 10), host)
    assert np.array_equal(g.sobol([0.0, 0.0], [1.0, 1.0], 1, 3), np.array([[0.5, 0.75, 0.25], [0.5, 0.25, 0.75]]))   # Sobol.jl's first points


def test_batched_lbfgs_against_restatement_and_options(bo):
    """b200bo_acquire_lbfgs: M box-bounded L-BFGS ascents in lock-step on the fused value+gradient launch, against the
    plain-Python restatement of the same iteration driven by the oracle's value and gradient; NLopt-style options (reference
    src/acquisition.jl:24-27): maxeval per start, ftol_rel stops early, iterates stay in the box, values never fall below the start."""
    from oracle import lbfgs_oracle as lo
    rng, o, g, X, y = make_pair(bo, "Mat52Ard", "MeanConst", 3, 220, seed=991)
    lb, ub = np.zeros(3), np.ones(3)
    M = 40
    X0 = rng.random((3, M)); X0[:, 0] = [0.0, 1.0, 0.5]                       # one start on the boundary
    tau = float(np.quantile(y, 0.8))
    for kind, par in (("EI", (tau,)), ("UCB", (2.0,))):
        f0 = g.acquire(kind, par, X0)["values"]
        r = g.acquire_lbfgs(kind, par, X0, lb, ub, maxeval=60, ftol_rel=1e-10)
        assert np.all(r["X"] >= 0.0) and np.all(r["X"] <= 1.0) and np.all(r["evals"] <= 60) and np.all(r["evals"] >= 1)
        assert np.all(r["values"] >= f0 - 1e-15 * np.abs(f0))
        assert np.mean(r["values"] > f0 + 1e-9 * np.abs(f0)) > 0.7                       # the ascent actually climbs
        chk = g.acquire(kind, par, r["X"])["values"]
        assert close(chk, r["values"], 1e-9)                                             # reported value = value at the reported point
        assert r["best_index"] == orc.first_strict_argmax_np(r["values"]) and np.array_equal(r["best_x"], r["X"][:, r["best_index"]])
        agree = 0
        for j in range(M):
            fg = lambda x: tuple(v[0] if np.ndim(v) == 1 else v[:, 0] for v in orc.acq_grad(o, kind, par, x.reshape(3, 1)))
            ro = lo.maximize(fg, X0[:, j], lb, ub, maxeval=60, ftol_rel=1e-10)
            agree += abs(ro.f - r["values"][j]) <= 1e-6 * abs(ro.f) + 1e-10
        assert agree >= 0.8 * M                                                          # same trajectories up to accept/reject ties
        few = g.acquire_lbfgs(kind, par, X0, lb, ub, maxeval=3)
        assert np.all(few["evals"] <= 3) and np.all(few["values"] <= r["values"] + 1e-9 * np.abs(r["values"]) + 1e-12)
    with pytest.raises(bo._lib.B200BOError):
        g.acquire_lbfgs("TS", (), X0, lb, ub)                                            # derivative-free in the reference (acquisition.jl:7-9)


def test_map_fit_in_library_matches_scipy_on_the_oracle(bo):
    """b200bo_map_fit (reference optimizemodel!, src/models/gp.jl:54-77): box-bounded L-BFGS over the device mll + gradient,
    bounds as `noisebounds` / `kernbounds` (gp.jl:65-68), against SciPy's L-BFGS-B on the oracle's objective from the same start."""
    from scipy.optimize import minimize
    rng, o, g, X, y = make_pair(bo, "SEArd", "MeanConst", 2, 160, seed=515)
    th0 = g.get_params()
    sel = np.array([0, 2, 3, 4])                                                         # noise + kernel; the mean stays (domean = false)
    lbp, ubp = np.array([-4.0, -3.0, -3.0, -2.0]), np.array([1.0, 2.0, 2.0, 2.0])
    r = g.map_fit(th0[sel], lbp, ubp, noise=True, domean=False, kern=True, maxeval=200, ftol_rel=1e-12)

    def negf(x):
        th = th0.copy(); th[sel] = x
        f, gr = o.mll_dmll(th)
        return -f, -gr[sel]
    ref = minimize(negf, th0[sel], jac=True, method="L-BFGS-B", bounds=list(zip(lbp, ubp)), options=dict(ftol=1e-14, gtol=1e-9, maxfun=400))
    assert r["status"] in (3, 4) and r["evals"] <= 200
    assert r["mll"] >= -ref.fun - 1e-6 * abs(ref.fun)                                    # at least as good an optimum
    assert np.all(r["theta"] >= lbp) and np.all(r["theta"] <= ubp)
    if abs(r["mll"] + ref.fun) <= 1e-7 * abs(ref.fun):
        assert np.allclose(r["theta"], ref.x, atol=2e-3)
    th = g.get_params()
    assert np.array_equal(th[sel], r["theta"]) and th[1] == th0[1]                       # the model sits at the optimum, mean untouched
    assert abs(g.mll - r["mll"]) <= 1e-10 * abs(r["mll"])
    three = g.map_fit(np.stack([th0[sel], th0[sel] + 0.3, th0[sel] - 0.3], axis=1), lbp, ubp, noise=True, domean=False, kern=True, maxeval=60)
    assert three["mll"] >= r["mll"] - 1e-6 * abs(r["mll"])                               # restarts in lock-step never do worse


def test_normal_priors_in_the_map_target(bo):
    """EXT set_priors! (reference src/models/gp.jl:30-35): Normal priors on chosen parameters enter the MAP target and its gradient; the
    fixed parameters' priors are constants of the sweep; clearing restores the flat target."""
    rng, o, g, X, y = make_pair(bo, "Mat32Ard", "MeanConst", 3, 140, seed=707)
    th = g.get_params()
    pri = [(-2.0, 0.5), None, (0.0, 1.0), None, (-0.3, 0.2), (0.1, 2.0)]
    m0, d0 = g.mll_sweep(np.stack([th, th + 0.1], axis=1))
    g.set_priors(pri); o.priors = pri
    m1, d1 = g.mll_sweep(np.stack([th, th + 0.1], axis=1))
    for k, t in enumerate((th, th + 0.1)):
        fo, go = o.mll_dmll(t)
        assert abs(m1[k] - fo) <= 1e-10 * abs(fo) and relmax(d1[:, k], go) < 1e-8
    assert np.all(m1 != m0)
    mk, dk = g.mll_sweep(np.stack([th[[0, 2, 3, 4, 5]]], axis=1), domean=False)          # beta fixed: its (flat) prior drops out
    o.set_params(th)                                                                     # the oracle is stateful: back to the model's beta
    fo, go = o.mll_dmll(th[[0, 2, 3, 4, 5]], domean=False)
    assert abs(mk[0] - fo) <= 1e-10 * abs(fo) and relmax(dk[:, 0], go) < 1e-8
    r = g.map_fit(th[[0, 2, 3, 4, 5]], np.full(5, -4.0), np.full(5, 3.0), domean=False, maxeval=80)
    assert r["mll"] >= mk[0] - 1e-9 * abs(mk[0])
    g.set_priors(None)
    m2, _ = g.mll_sweep(np.stack([th, th + 0.1], axis=1))
    assert np.array_equal(m2, m0)


@pytest.mark.parametrize("kern,N0,m", [("SEArd", 1000, 16), ("Mat52Ard", 250, 9), ("SEArd", 120, 16), ("Mat32Iso", 128, 5)])
def test_blocked_append_from_the_inverse_factor(bo, kern, N0, m):
    """A batch of m points appended right after an acquisition (the BO loop's order: acquire_max, then update!, reference
    src/BayesianOptimization.jl:185-196): the new factor rows come from ONE dense product with W = L^-1 (U12 = U11^-T A12 of EXT
    ElasticPDMats append!), the m x m block from a small Cholesky; batches that cross a 128-block boundary finish point by point.
    Against a full refactor and the oracle, then again after a second acquisition + append."""
    import time
    D = 4
    rng = np.random.default_rng(N0 + m)
    X = rng.random((D, N0 + 2 * m)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N0 + 2 * m)
    ll = np.full(1 if kern.endswith("Iso") else D, np.log(0.5))
    mk = lambda: bo.B200GPE(D, mean=bo.MeanConst(0.2), kernel=bo.gp._Kernel(kern, ll, 0.1), logNoise=-2.0, capacity=N0 + 300)
    g = mk(); g.fit(X[:, :N0], y[:N0])
    Xs = rng.random((D, 500))
    for rnd in range(2):
        n1 = N0 + (rnd + 1) * m
        g.acquire("EI", (0.3,), Xs)                                   # leaves W = L^-1 of the current factor behind
        t0 = time.perf_counter(); bo.update(g, X[:, n1 - m:n1], y[n1 - m:n1]); dt = time.perf_counter() - t0
        ref = mk(); ref.fit(X[:, :n1], y[:n1])
        o = orc.GPOracle(D, kern, "MeanConst", ll=ll, lsigma=0.1, lognoise=-2.0, beta=0.2).fit(X[:, :n1], y[:n1])
        assert bo.dims(g) == (D, n1)
        assert relmax(g.factor, o.U) < 1e-11 and relmax(g.alpha, o.alpha) < 1e-9 and abs(g.mll - o.mll) < 1e-10 * abs(o.mll)
        assert relmax(g.factor, ref.factor) < 1e-12
        r, r0 = g.acquire("EI", (0.3,), Xs, want_grad=True, want_mu_var=True), ref.acquire("EI", (0.3,), Xs, want_grad=True, want_mu_var=True)
        assert close(r["mu"], r0["mu"], 1e-9) and close(r["var"], r0["var"], 1e-8, 1e-13) and r["best_index"] == r0["best_index"]
    print(f"append of {m} points at N={N0 + m}: {dt * 1e3:.3f} ms")


# ------------------------------------------------------------------------------------------------------------
# joint posterior sample: myrand(model, X::Matrix) -> EXT rand(gp, X) (src/models/gp.jl:7)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel,N,M", [("SEArd", 200, 40), ("Mat52Ard", 700, 300), ("SEArd", 0, 50)])
def test_joint_posterior_sample_matches_restatement(kernel, N, M):
    import b200bo
    rng = np.random.default_rng(N + M)
    D = 3
    X = rng.random((D, N)); y = np.sin(4 * X.sum(axis=0)) + 0.05 * rng.normal(size=N)
    ll = np.full(D, -1.6)                      # short length-scale: the posterior covariance of M scattered points is well conditioned
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.3), kernel=getattr(b200bo, kernel)(ll, 0.2), logNoise=-1.5, capacity=max(N, 128))
    o = orc.GPOracle(D, kernel, "MeanConst", ll=ll, lsigma=0.2, lognoise=-1.5, beta=0.3)
    if N:
        g.fit(X, y)
    o.fit(X, y)
    Xs = rng.random((D, M))
    seed, off = 77, 1000
    r = g.rand_joint(Xs, seed=seed, idx_offset=off)
    eps = orc.philox_normal(seed, off + np.arange(M))
    want, tries = o.rand_joint(Xs, eps)
    mu, S = o.posterior_cov(Xs)
    assert r["tries"] == tries == 0
    assert np.all(np.abs(r["mu"] - mu) <= 1e-6 * np.abs(mu) + 1e-12)
    scale = np.sqrt(np.diag(S).max())
    assert np.abs(r["sample"] - want).max() <= 1e-7 * scale       # the Schur complement inside the augmented factor == chol(K** - V'V)
    # matrix form of the generic function (gp.jl:7) is the joint draw; the vector form (gp.jl:6) stays an independent draw
    assert np.array_equal(b200bo.myrand(g, Xs, seed=seed, idx_offset=off), r["sample"])
    one = b200bo.myrand(g, Xs[:, 0], seed=seed, idx_offset=off)
    m1, v1 = o.predict(Xs[:, :1])
    assert abs(one - (m1[0] + np.sqrt(v1[0]) * eps[0])) <= 1e-9 * scale


def test_joint_posterior_sample_statistics_and_make_posdef():
    """duplicate columns make Sigma_post exactly singular: make_posdef!'s jitter rule must kick in and the draws at the duplicates agree to
    sqrt(jitter); over many seeds the empirical covariance approaches Sigma_post"""
    import b200bo
    rng = np.random.default_rng(9)
    D, N, M = 2, 60, 12
    X = rng.random((D, N)); y = np.cos(3 * X[0]) * X[1]
    g = b200bo.B200GPE(D, mean=b200bo.MeanZero(), kernel=b200bo.SEArd(np.full(D, -1.0), 0.0), logNoise=-2.0, capacity=128)
    g.fit(X, y)
    o = orc.GPOracle(D, "SEArd", "MeanZero", ll=np.full(D, -1.0), lsigma=0.0, lognoise=-2.0).fit(X, y)
    Xs = rng.random((D, M)) * 1.5
    mu, S = o.posterior_cov(Xs)
    draws = np.stack([g.rand_joint(Xs, seed=s)["sample"] for s in range(1500)])
    emp = np.cov(draws.T)
    assert np.abs(draws.mean(axis=0) - mu).max() < 5 * np.sqrt(np.diag(S).max() / 1500)
    assert np.abs(emp - S).max() < 0.15 * np.diag(S).max()
    dup = np.concatenate([Xs, Xs[:, :3]], axis=1)
    r = g.rand_joint(dup, seed=3)
    sd = np.sqrt(np.diag(S)[:3])
    assert r["tries"] <= 10 and np.all(np.abs(r["sample"][M:] - r["sample"][:3]) <= 0.05 * sd + 1e-3)
    # 150 numerically coincident points: 149 pivots of rounding noise cannot all come out positive -> the jitter rule must act
    cl = np.array([[0.4], [0.6]]) + 1e-10 * rng.random((D, 150))
    rc = g.rand_joint(cl, seed=4)
    _, tries_o = o.rand_joint(cl, orc.philox_normal(4, np.arange(150)))
    assert 1 <= rc["tries"] <= 10 and tries_o >= 1
    m0, v0 = o.predict(cl[:, :1])
    assert np.all(np.abs(rc["sample"] - rc["sample"][0]) <= 0.02 * np.sqrt(v0[0]) + 1e-3)       # one draw, shared by the whole cluster


def test_pinned_pipelined_transfers_equal_pageable_path(bo):
    """host-pointer b200bo_acquire: with pinned candidate / output buffers the library moves the data chunk by chunk on its stream lanes
    (under the kernels of the neighbouring chunk); pageable buffers are copied in one piece.  Same bits either way."""
    torch = pytest.importorskip("torch")
    _, o, g, X, y = make_pair(bo, "Mat52Ard", "MeanConst", 5, 600, seed=21)
    rng = np.random.default_rng(22)
    M = 40000                                   # three chunks at N = 640 (64 MB of k* slices each)
    Xs = np.asfortranarray(rng.random((5, M)))
    tau = float(y.max())
    r1 = g.acquire("EI", (tau,), Xs, want_grad=True)
    Xp = torch.from_numpy(np.ascontiguousarray(Xs.T)).pin_memory().numpy().T
    vp = torch.empty(M, dtype=torch.float64).pin_memory().numpy()
    gp_ = torch.empty((M, 5), dtype=torch.float64).pin_memory().numpy().T
    r2 = g.acquire("EI", (tau,), Xp, want_grad=True, values_out=vp, grad_out=gp_)
    assert r2["values"] is vp and r2["grad"] is gp_
    assert np.array_equal(r1["values"], r2["values"]) and np.array_equal(r1["grad"], r2["grad"])
    assert r1["best_index"] == r2["best_index"] and r1["best_value"] == r2["best_value"] and np.array_equal(r1["best_x"], r2["best_x"])
    g.set_acq_engine(0)                         # the DMMA engine has no chunk lanes: one copy in, one out
    r3 = g.acquire("EI", (tau,), Xp, want_grad=True, values_out=vp, grad_out=gp_)
    g.set_acq_engine(1)
    assert r3["best_index"] == r1["best_index"] and np.abs(r3["values"] - r1["values"]).max() <= 1e-9 * np.abs(r1["values"]).max()


def test_cholesky_schedules_agree_and_graph_replay_is_bit_identical(bo):
    """the look-ahead schedule (fused cluster head) against the in-order schedule of round 1, and its CUDA-graph replay against eager launches"""
    rng = np.random.default_rng(31)
    D, N = 4, 1700                              # 14 panels: three full outer panels + a ragged one, tcgen05 far / near updates included
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    g = bo.B200GPE(D, mean=bo.MeanConst(0.1), kernel=bo.Mat32Ard(np.full(D, -0.6), 0.2), logNoise=-2.0, capacity=N)
    out = {}
    for name, (sched, graph) in dict(inorder=(0, 0), eager=(1, 0), tile_heads=(2, 0), graph=(1, 2)).items():
        g.set_knob("chol_sched", sched); g.set_knob("chol_graph", graph)
        fs = []
        for _ in range(3):                      # graph mode (knob 2): eager first sight, capture on the second, replay on the third
            g.fit(X, y); fs.append((g.factor, g.alpha, g.mll))
        assert all(np.array_equal(fs[0][0], f[0]) and np.array_equal(fs[0][1], f[1]) and fs[0][2] == f[2] for f in fs[1:]), name
        out[name] = fs[-1]
    assert np.array_equal(out["eager"][0], out["graph"][0]) and np.array_equal(out["eager"][1], out["graph"][1])
    for name in ("eager", "tile_heads"):
        assert np.abs(out[name][0] - out["inorder"][0]).max() <= 1e-12 * np.abs(out["inorder"][0]).max()
        assert abs(out[name][2] - out["inorder"][2]) <= 1e-10 * abs(out["inorder"][2])
    o = orc.GPOracle(D, "Mat32Ard", "MeanConst", ll=np.full(D, -0.6), lsigma=0.2, lognoise=-2.0, beta=0.1).fit(X, y)
    assert relmax(out["graph"][1], o.alpha) <= 1e-9 and abs(out["graph"][2] - o.mll) <= 1e-10 * abs(o.mll)


def test_replayed_factorisation_graph_follows_data_and_mean_changes(bo):
    """the CUDA graph of the factorisation is keyed on the panel count and the buffers; N (inside the same 128-padding) and the mean
    parameter change underneath it -- the right-hand side y - m must follow (it is computed outside the graph)"""
    rng = np.random.default_rng(41)
    D, N = 3, 1000
    X = rng.random((D, N + 10)); y = np.cos(2 * X.sum(0)) + 0.05 * rng.standard_normal(N + 10) + 1.5
    ll = np.full(D, -0.8)
    g = bo.B200GPE(D, mean=bo.MeanConst(0.0), kernel=bo.SEArd(ll, 0.0), logNoise=-2.0, capacity=1024)
    g.set_knob("chol_graph", 2)                  # capture at the second consecutive factorisation of a shape (default: the sixth)
    for _ in range(3):
        g.fit(X[:, :N], y[:N])                   # eager, capture, replay
    for beta, n in ((0.0, N), (1.5, N), (1.5, N + 7), (-0.4, N + 10), (0.9, N - 3)):
        th = g.get_params(); th[1] = beta
        g.set_params(th)
        g.fit(X[:, :n], y[:n])
        o = orc.GPOracle(D, "SEArd", "MeanConst", ll=ll, lsigma=0.0, lognoise=-2.0, beta=beta).fit(X[:, :n], y[:n])
        assert relmax(g.alpha, o.alpha) <= 1e-9, (beta, n)
        assert abs(g.mll - o.mll) <= 1e-10 * abs(o.mll), (beta, n)
    # the MAP objective on a worker model (one setting in flight: the worker replays its own graph) with the mean parameter swept
    Theta = np.stack([np.concatenate([[-2.0, b], ll, [0.0]]) for b in (-1.0, 0.0, 0.7, 1.5, 2.2)], axis=1)
    for _ in range(3):
        mll = np.array([g.mll_sweep(Theta[:, [s]], want_grad=False)[0][0] for s in range(Theta.shape[1])])
    n = N - 3
    want = [orc.GPOracle(D, "SEArd", "MeanConst", ll=ll, lsigma=0.0, lognoise=-2.0, beta=b).fit(X[:, :n], y[:n]).mll for b in Theta[1]]
    assert np.all(np.abs(mll - np.array(want)) <= 1e-10 * np.abs(want))
