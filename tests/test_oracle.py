"""Pins the CPU restatement (oracle/) -- the checker every GPU parity test relies on.

What the reference's own tests pin at this boundary is restated here: test/acquisition.jl:11-12
(1-point GP, arg-max of the posterior mean at 1.0), test/acquisitionfunctions.jl:8-10 (batched == per point),
test/warmstart.jl:64 (tau == max y).  Numeric values of mu / s2 / a / grad / mll are NOT pinned by the reference
("parity unpinned"): they are anchored by analytic closed forms, 50-digit mpmath evaluation, finite differences, a
published Philox known-answer vector, the committed golden fixtures and oracle.c vs LAPACK agreement.
"""
import glob
import math
import os

import mpmath as mp
import numpy as np
import pytest

from oracle import gp_oracle as orc
from oracle.c_oracle import COracle


def test_one_point_gp_known_answer():
    # test/acquisition.jl:1-12: GPE([1.0],[2.0],MeanZero(),SEIso(1.0,0.0)), default logNoise -2
    gp = orc.GPOracle(1, "SEIso", "MeanZero", ll=[1.0], lsigma=0.0, lognoise=-2.0).fit(np.array([[1.0]]), np.array([2.0]))
    mu, var = gp.predict(np.array([[1.0]]))
    den = 1.0 + math.exp(-4.0) + orc.EPS
    assert mu[0] == pytest.approx(2.0 / den, rel=1e-15)
    assert var[0] == pytest.approx(1.0 - 1.0 / den, rel=1e-13)
    xs = np.linspace(-5, 5, 4001)[None, :]
    m, _ = gp.predict(xs)
    assert xs[0, np.argmax(m)] == pytest.approx(1.0, abs=1e-12)
    x = 2.5
    assert gp.predict(np.array([[x]]))[0][0] == pytest.approx(2.0 * math.exp(-(x - 1) ** 2 / (2 * math.e ** 2)) / den, rel=1e-14)


def test_two_point_closed_form():
    ell, sf2, noise = 0.7, math.exp(0.4), math.exp(-3.0) + orc.EPS
    X = np.array([[0.0, 1.0]]); y = np.array([0.5, -0.2])
    gp = orc.GPOracle(1, "SEIso", "MeanConst", ll=[math.log(ell)], lsigma=0.2, lognoise=-1.5, beta=0.1).fit(X, y)
    k = lambda a, b: sf2 * math.exp(-0.5 * (a - b) ** 2 / ell ** 2)
    a_, b_ = sf2 + noise, k(0, 1)
    det = a_ * a_ - b_ * b_
    xs = 0.3
    ks = np.array([k(0, xs), k(1, xs)])
    Sinv = np.array([[a_, -b_], [-b_, a_]]) / det
    assert gp.predict(np.array([[xs]]))[0][0] == pytest.approx(0.1 + ks @ Sinv @ (y - 0.1), rel=1e-13)
    assert gp.predict(np.array([[xs]]))[1][0] == pytest.approx(sf2 - ks @ Sinv @ ks, rel=1e-12)
    assert gp.mll == pytest.approx(-0.5 * ((y - 0.1) @ Sinv @ (y - 0.1) + math.log(det) + 2 * math.log(2 * math.pi)), rel=1e-13)


@pytest.mark.parametrize("kern", ["SEArd", "Mat12Iso", "Mat32Ard", "Mat52Ard"])
def test_posterior_against_mpmath(kern):
    """50-digit evaluation of mu and s2 (incl. the k** - v'v cancellation on top of a training point)."""
    mp.mp.dps = 50
    rng = np.random.default_rng(7)
    D, N = 2, 12
    X = rng.random((D, N)); y = rng.standard_normal(N)
    ll = np.array([-0.4]) if kern.endswith("Iso") else np.array([-0.4, -0.1])
    gp = orc.GPOracle(D, kern, "MeanConst", ll=ll, lsigma=0.1, lognoise=-2.0, beta=0.2).fit(X, y)
    ie = [mp.e ** (-mp.mpf(float(v))) for v in (ll if ll.size == D else [ll[0]] * D)]
    sf2 = mp.e ** (2 * mp.mpf(0.1))

    def kf(a, b):
        r2 = sum(((mp.mpf(float(a[d])) - mp.mpf(float(b[d]))) * ie[d]) ** 2 for d in range(D))
        r = mp.sqrt(r2)
        if kern.startswith("SE"):
            return sf2 * mp.e ** (-r2 / 2)
        if kern.startswith("Mat12"):
            return sf2 * mp.e ** (-r)
        if kern.startswith("Mat32"):
            s = mp.sqrt(3) * r
            return sf2 * (1 + s) * mp.e ** (-s)
        s = mp.sqrt(5) * r
        return sf2 * (1 + s + s * s / 3) * mp.e ** (-s)

    S = mp.matrix(N, N)
    for i in range(N):
        for j in range(N):
            S[i, j] = kf(X[:, i], X[:, j]) + ((mp.e ** (2 * mp.mpf(-2.0)) + mp.mpf(orc.EPS)) if i == j else 0)
    r = mp.matrix([mp.mpf(float(v)) - mp.mpf(0.2) for v in y])
    alpha = mp.lu_solve(S, r)
    for xs in (X[:, 3], rng.random(D)):
        ks = mp.matrix([kf(X[:, i], xs) for i in range(N)])
        mu = mp.mpf(0.2) + sum(ks[i] * alpha[i] for i in range(N))
        w = mp.lu_solve(S, ks)
        s2 = sf2 - sum(ks[i] * w[i] for i in range(N))
        m_o, v_o = gp.predict(xs.reshape(D, 1))
        assert float(abs(m_o[0] - mu) / abs(mu)) < 1e-11
        assert float(abs(v_o[0] - s2) / abs(s2)) < 1e-9


def test_functors_as_coded_against_mpmath():
    """EI = Delta*Phi(z) + phi(z) (quirk 1, NOT textbook), PI via erf, exact-zero branches (quirks 2-3)."""
    mp.mp.dps = 40
    for mu, tau, s2 in [(0.3, 0.5, 0.01), (1.2, 0.1, 0.5), (-0.4, 0.0, 2.0), (0.7, 0.7, 1e-6)]:
        d = mp.mpf(mu) - mp.mpf(tau); s = mp.sqrt(mp.mpf(s2)); z = d / s
        Phi = (1 + mp.erf(z / mp.sqrt(2))) / 2; phi = mp.e ** (-z * z / 2) / mp.sqrt(2 * mp.pi)
        assert orc.acq_value("EI", (tau,), mu, s2) == pytest.approx(float(d * Phi + phi), rel=1e-12)
        assert orc.acq_value("PI", (tau,), mu, s2) == pytest.approx(float(Phi), rel=1e-12)
    assert orc.acq_value("EI", (0.5,), 0.3, 0.01) == pytest.approx(0.049441, abs=2e-6)     # SURVEY 0.4-1 scratch value
    assert orc.acq_value("EI", (0.5,), 0.7, 0.0) == pytest.approx(0.2) and orc.acq_value("EI", (0.5,), 0.3, 0.0) == 0.0
    assert orc.acq_value("PI", (0.5,), 0.7, 0.0) == 1.0 and orc.acq_value("PI", (0.5,), 0.5, 0.0) == 0.0
    assert orc.acq_value("UCB", (2.0,), 0.3, 0.25) == pytest.approx(1.3)
    assert orc.acq_value("MI", (1.5, 0.2), 0.3, 0.05) == pytest.approx(0.3 + 1.5 * (math.sqrt(0.25) - math.sqrt(0.2)))
    assert orc.brochu_beta(6, 2048) == pytest.approx(math.sqrt(2 * math.log(2048 ** 5 * math.pi ** 2 / 0.3)))
    assert orc.brochu_beta(2, 0) == orc.brochu_beta(2, 1)


@pytest.mark.parametrize("kern", orc.KERNELS)
def test_gradients_against_finite_differences(kern):
    rng = np.random.default_rng(11)
    D, N = 3, 40
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    gp = orc.GPOracle(D, kern, "MeanConst", ll=rng.normal(-0.5, 0.2, 1 if kern.endswith("Iso") else D), lsigma=0.1,
                      lognoise=-1.5, beta=0.3).fit(X, y)
    Xs = rng.random((D, 5))
    h = 1e-6
    for acq, par in [("EI", (np.median(y),)), ("PI", (np.median(y),)), ("UCB", (2.0,)), ("MI", (1.0, 0.3)), ("MaxMean", ())]:
        a, g = orc.acq_grad(gp, acq, par, Xs)
        fd = np.zeros_like(g)
        for d in range(D):
            Xp, Xm = Xs.copy(), Xs.copy(); Xp[d] += h; Xm[d] -= h
            fd[d] = (orc.acq_value(acq, par, *gp.predict(Xp)) - orc.acq_value(acq, par, *gp.predict(Xm))) / (2 * h)
        assert np.abs(g - fd).max() / np.abs(fd).max() < 1e-6
    th = gp.get_params()
    f0, g0 = gp.mll_dmll(th)
    fd = np.array([(gp.mll_dmll(th + h * e)[0] - gp.mll_dmll(th - h * e)[0]) / (2 * h) for e in np.eye(th.size)])
    assert np.abs(g0 - fd).max() / np.abs(fd).max() < 1e-6


def test_batched_equals_per_point_and_selection_rule():
    # test/acquisitionfunctions.jl:1-12 shape: GPE(rand(3,4), rand(4), MeanZero(), SEIso(0,0)), x = rand(3,2)
    rng = np.random.default_rng(3)
    gp = orc.GPOracle(3, "SEIso", "MeanZero", ll=[0.0], lsigma=0.0).fit(rng.random((3, 4)), rng.random(4))
    x = rng.random((3, 2))
    mu, var = gp.predict_column_loop(x)
    m1, v1 = gp.predict_column_loop(x[:, :1])
    assert mu[0] == m1[0] and var[0] == v1[0]
    for kind, par in [("PI", (0.5,)), ("EI", (0.5,)), ("UCB", (1.0,)), ("MI", (1.0, 0.0))]:
        assert len(orc.acq_value(kind, par, mu, var)) == 2
        assert orc.acq_value(kind, par, mu, var)[0] == orc.acq_value(kind, par, m1, v1)[0]
    # acquire_max keeps the FIRST strict maximum, NaN never wins, nothing wins -> -1 (acquisition.jl:55-66)
    assert orc.first_strict_argmax([1.0, 3.0, 3.0, np.nan, 2.0]) == 1 == orc.first_strict_argmax_np([1.0, 3.0, 3.0, np.nan, 2.0])
    assert orc.first_strict_argmax([np.nan, -np.inf]) == -1 == orc.first_strict_argmax_np([np.nan, -np.inf])
    assert orc.first_strict_argmax([]) == -1 == orc.first_strict_argmax_np([])


def test_setparams_semantics():
    rng = np.random.default_rng(5)
    X = rng.random((2, 10)); y = rng.standard_normal(10)
    gp = orc.GPOracle(2, "SEArd", "MeanConst", ll=[0.0, 0.0], lsigma=0.0).fit(X, y)
    assert orc.maxy(gp) == y.max()                               # test/warmstart.jl:64 (tau = max(maxy, tau))
    assert orc.maxy(orc.GPOracle(2)) == -math.inf
    g1 = orc.mi_gamma_update(gp, 0.0)
    assert g1 == pytest.approx(gp.predict(X[:, -1:])[1][0])
    assert orc.mi_gamma_update(orc.GPOracle(2), 3.0) == 0.0


def test_jitter_rule_make_posdef():
    X = np.linspace(0.0, 1.0, 60)[None, :]; y = np.sin(4 * X[0])   # smooth kernel, l = e^3: numerically rank-deficient
    gp = orc.GPOracle(1, "SEIso", "MeanZero", ll=[3.0], lsigma=0.0, lognoise=-30.0).fit(X, y)
    assert gp.jitter_tries == 1                                  # one 1e-6 tr(Sigma)/n bump makes it factorisable
    assert np.all(np.isfinite(gp.alpha))


def test_philox_known_answer_and_moments():
    # Random123 kat_vectors, philox4x32 10 rounds
    z = orc.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0]
    assert [int(v) for v in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = orc.philox4x32_10(np.full((1, 4), 0xFFFFFFFF, np.uint32), np.full((1, 2), 0xFFFFFFFF, np.uint32))[0]
    assert [int(v) for v in f] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    p = orc.philox4x32_10(np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32),
                          np.array([[0xa4093822, 0x299f31d0]], np.uint32))[0]
    assert [int(v) for v in p] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    e = orc.philox_normal(50, np.arange(200000))
    assert abs(e.mean()) < 0.01 and abs(e.std() - 1.0) < 0.01
    assert np.array_equal(orc.philox_normal(50, np.arange(100, 110)), e[100:110])      # keyed by global index


def test_latin_hypercube_is_stratified():
    rng = np.random.default_rng(0)
    lb, ub = np.array([-5.0, 0.0, 2.0]), np.array([10.0, 15.0, 2.0])
    S = orc.latin_hypercube_sampling(lb, ub, 64, rng)
    assert S.shape == (3, 64)
    for d in range(2):
        strata = np.floor((S[d] - lb[d]) / ((ub[d] - lb[d]) / 64)).astype(int)
        assert sorted(strata) == list(range(64))
    assert np.all(S[2] == 2.0)
    with pytest.raises(ValueError):
        orc.latin_hypercube_sampling([0, 0], [1], 4, rng)
    with pytest.raises(ValueError):
        orc.latin_hypercube_sampling([1.0], [0.0], 4, rng)


def test_golden_fixtures_regression(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "*.npz")))
    assert len(files) >= 6
    for f in files:
        z = np.load(f)
        gp = orc.GPOracle(int(z["D"]), str(z["kernel"]), str(z["mean"]))
        gp.set_params(z["theta"])
        gp.fit(z["X"], z["y"])
        assert np.allclose(gp.alpha, z["alpha"], rtol=1e-9, atol=1e-12)
        assert gp.mll == pytest.approx(float(z["mll"]), rel=1e-12)
        mu, var = gp.predict(z["Xs"])
        assert np.allclose(mu, z["mu"], rtol=1e-10, atol=1e-13) and np.allclose(var, z["var"], rtol=1e-8, atol=1e-13)
        for k in ("EI", "PI", "UCB", "MI", "MaxMean"):
            a, g = orc.acq_grad(gp, k, tuple(z[f"{k}_params"]), z["Xs"])
            assert np.allclose(a, z[f"{k}_values"], rtol=1e-8, atol=1e-13)
            assert orc.first_strict_argmax_np(a) == int(z[f"{k}_best"])
        f2, g2 = gp.mll_dmll(z["theta2"])
        assert f2 == pytest.approx(float(z["mll2"]), rel=1e-12) and np.allclose(g2, z["dmll2"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("kern", ["SEArd", "Mat52Iso", "Mat12Ard", "Mat32Ard"])
def test_c_restatement_matches_lapack_oracle(kern):
    rng = np.random.default_rng(0)
    D, N, M = 5, 200, 40
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    gp = orc.GPOracle(D, kern, "MeanConst", ll=rng.normal(-0.3, 0.1, 1 if kern.endswith("Iso") else D), lsigma=0.1, lognoise=-2,
                      beta=0.2).fit(X, y)
    co = COracle(gp, refit_in_c=True)
    assert np.abs(co.U - gp.U).max() < 1e-12 and co.mll == pytest.approx(gp.mll, rel=1e-12)
    assert np.allclose(co.alpha, gp.alpha, rtol=1e-9, atol=1e-11)
    Xs = rng.random((D, M))
    mu, var = gp.predict(Xs)
    for kind, par in [("EI", (np.quantile(y, .9),)), ("PI", (np.quantile(y, .9),)), ("UCB", (3.0,)), ("MI", (1.0, .2)), ("MaxMean", ()),
                      ("TS", ())]:
        r = co.acquire(kind, par, Xs, seed=5, idx_offset=7, want_grad=kind != "TS", nthreads=2)
        if kind == "TS":
            a = orc.acq_value(kind, par, mu, var, eps=orc.philox_normal(5, 7 + np.arange(M)))
        else:
            a, g = orc.acq_grad(gp, kind, par, Xs)
            assert np.abs(r["grad"] - g).max() / np.abs(g).max() < 1e-9
        assert np.allclose(r["values"], a, rtol=1e-9, atol=1e-13)
        assert r["best_index"] - 7 == orc.first_strict_argmax_np(a)


def test_device_lhs_restatement_is_a_latin_hypercube():
    lb, ub = np.array([-1.0, 3.0]), np.array([2.0, 4.0])
    for n, seed in [(1, 0), (2, 1), (33, 2), (512, 3)]:
        X = orc.lhs_device(lb, ub, n, seed)
        for d in range(2):
            assert sorted(np.floor((X[d] - lb[d]) / ((ub[d] - lb[d]) / n)).astype(int)) == list(range(n))
    assert np.array_equal(orc.lhs_device(lb, ub, 100, 9, offset=40, n_local=20), orc.lhs_device(lb, ub, 100, 9)[:, 40:60])
    assert not np.array_equal(orc.lhs_device(lb, ub, 100, 9), orc.lhs_device(lb, ub, 100, 10))


def test_ascent_restatement_improves():
    rng = np.random.default_rng(2)
    X = rng.random((2, 30)); y = np.sin(4 * X[0]) + np.cos(3 * X[1])
    gp = orc.GPOracle(2, "SEArd", "MeanZero", ll=[-1.0, -1.0], lsigma=0.0).fit(X, y)
    X0 = rng.random((2, 20))
    Xb, Fb = orc.ascent(gp, "MaxMean", (), X0, np.zeros(2), np.ones(2), steps=25)
    f0 = orc.acq_value("MaxMean", (), *gp.predict(X0))
    assert np.all(Fb >= f0) and Fb.mean() > f0.mean() + 0.05 and np.all((Xb >= 0) & (Xb <= 1))


def test_int8_slicing_is_exact_and_fp64_accurate():
    """The digit rule of the tcgen05 trailing update (csrc/syrk_i8.cu, restated in oracle/ozaki.py): digits fit int8, the 7-slice
    representation is exact to 54 bits, and the 28-product reconstruction matches a long-double product to FP64 round-off."""
    from oracle import ozaki
    rng = np.random.default_rng(0)
    A = rng.standard_normal((96, 512)) * np.exp(rng.uniform(-12, 3, (96, 1)))       # rows of very different scale
    A[:, :40] *= 1e-7                                                               # and wide dynamic range inside a row
    A[5] = 0.0
    A[6, 17] = 1.0; A[6, 18] = -0.49999999999999994                                 # half-way digits
    D, sc = ozaki.slice_rows(A)
    assert D[0].min() >= -64 and D[0].max() <= 64 and D[1:].min() >= -128 and D[1:].max() <= 127
    rowmax = np.abs(A).max(axis=1, keepdims=True)
    assert np.all(np.abs(ozaki.reconstruct(D, sc) - A) <= 2.0 ** -53 * rowmax)
    B = rng.standard_normal((64, 512)) * np.exp(rng.uniform(-6, 6, (64, 1)))
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(float)
    got = ozaki.product(A, B)
    bound = 2.0 ** -49 * np.sqrt(512) * rowmax * np.abs(B).max(axis=1)[None, :]
    assert np.all(np.abs(got - ref) <= bound)
    plain = A @ B.T                                                                  # ordinary FP64 product: same error class
    assert np.abs(got - ref).max() <= 8 * max(np.abs(plain - ref).max(), 1e-300) + bound.min()


def test_sliced_forward_solve_keeps_posterior_accuracy():
    """Design check for moving the acquisition kernel's off-diagonal products to int8 slices (tcgen05 findings):
    L sliced once with one scale per ROW over the whole row, the solve panel V sliced with the FIXED scale sigma_f (|v| <= sigma_f
    because ||v||^2 <= k**), diagonal-block solves in FP64.  The posterior variance -- the cancelling quantity -- must stay within
    the north_star tolerance by orders of magnitude."""
    from oracle import ozaki
    from scipy.linalg import solve_triangular
    rng = np.random.default_rng(5)
    D, N, M, nb = 6, 512, 48, 128
    X = rng.random((D, N)); y = -orc.hartmann6(X)
    o = orc.GPOracle(D, "Mat52Ard", "MeanConst", ll=np.zeros(D), lsigma=0.3, lognoise=-2.0, beta=0.0).fit(X, y)
    Xs = rng.random((D, M)); Xs[:, 0] = X[:, 7]; Xs[:, 1] = X[:, 300] + 1e-7           # on / next to training points: s2 cancels
    L = o.U.T.copy()
    ks = o.cov(o.X, Xs)                                                               # N x M
    sf = np.exp(o.lsigma)
    eV = int(np.frexp(sf * (1 + 1e-12))[1])                                           # |v| <= sigma_f < 2^eV
    eL = np.frexp(np.abs(L).max(axis=1))[1]                                           # one exponent per row of L, whole row
    V = np.zeros((N, M))
    for i in range(0, N, nb):
        R = ks[i:i + nb].copy()
        if i > 0:
            R -= ozaki.product(L[i:i + nb, :i], V[:i].T.copy(), ea=eL[i:i + nb], eb=eV)
        V[i:i + nb] = solve_triangular(L[i:i + nb, i:i + nb], R, lower=True)
    Vref = solve_triangular(L.astype(np.longdouble).astype(float), ks, lower=True)
    var = np.maximum(sf ** 2 - np.sum(V * V, axis=0), 0.0)
    mo, vo = o.predict(Xs)
    vref = np.maximum(sf ** 2 - np.sum(Vref * Vref, axis=0), 0.0)
    assert np.abs(V - Vref).max() <= 1e-13 * sf
    assert np.all(np.abs(var - vo) <= 1e-9 * np.abs(vo) + 1e-13) and np.all(np.abs(vref - vo) <= 1e-9 * np.abs(vo) + 1e-13)


def test_integer_digit_rule_of_the_tcgen05_acquisition_path():
    """csrc/umma.cuh i8_digits + csrc/acq_i8.cu gather_byte restated: digits fit int8 (top digit in [-64, 64]) and rebuild the 55-bit integer."""
    from oracle import ozaki
    rng = np.random.default_rng(5)
    q = np.concatenate([rng.integers(-2 ** 54, 2 ** 54, 4000), [0, 1, -1, 2 ** 54, -2 ** 54, 127, 128, -128, -129, 2 ** 47, 2 ** 48 - 1, -(2 ** 48)]])
    d = ozaki.int_digits(q)
    assert np.all(d[1:] >= -128) and np.all(d[1:] <= 127) and np.all(np.abs(d[0]) <= 64)
    back = sum(int(1) * d[s].astype(object) * (1 << (8 * (6 - s))) for s in range(7))
    assert np.all(back == q.astype(object))
    A = rng.standard_normal((16, 256)) * np.exp(rng.normal(0, 3, (16, 1)))
    D, sc = ozaki.slice_rows_int(A)
    rec = sum(D[s] * 2.0 ** (-8 * s) for s in range(6, -1, -1)) * sc[:, None]
    assert np.all(np.abs(rec - A) <= 2.0 ** -54 * np.abs(A).max(axis=1, keepdims=True))


def test_explicit_inverse_sliced_product_keeps_posterior_accuracy():
    """The arithmetic of K6 on tcgen05 (csrc/acq_i8.cu) on the CPU: sigma^2 from the explicit inverse factor through 7 x 7 int8 slices
    (28 exact products, anti-diagonal int32 accumulators) against the LAPACK triangular solve of the restated predict_f -- on random
    candidates, next to training points and ON training points (where k** - v'v cancels most)."""
    import scipy.linalg as sl
    from oracle import ozaki
    rng = np.random.default_rng(12)
    for kern, D, N, lognoise in (("Mat52Ard", 6, 512, -2.0), ("SEArd", 3, 384, -3.0), ("Mat12Ard", 2, 256, -2.0)):
        X = rng.random((D, N)); y = np.sin(3 * X.sum(0))
        o = orc.GPOracle(D, kern, "MeanZero", ll=np.full(D, -0.7), lsigma=0.3, lognoise=lognoise).fit(X, y)
        Xs = rng.random((D, 96)); Xs[:, :24] = X[:, :24]; Xs[:, 24:48] = X[:, 24:48] + 1e-7
        Ks = o.cov(Xs, o.X)                                                      # [cands][N]
        mo, vo = o.predict(Xs)
        got = ozaki.posterior_var_i8(o.U.T.copy(), Ks, o.sf2)
        assert np.all(np.abs(got - vo) <= 1e-9 * np.abs(vo) + 1e-13), kern
        assert np.max(np.abs(got - vo)) < 2e-13 * o.sf2, kern


def test_lbfgs_restatement_on_a_bounded_quadratic_and_options():
    """oracle/lbfgs_oracle.py (the restatement of csrc/lbfgs.cuh): converges to the constrained maximiser, honours maxeval / ftol / xtol
    the way NLopt counts them (reference src/acquisition.jl:24-27), never leaves the box, accepted values never decrease."""
    from oracle import lbfgs_oracle as lo
    rng = np.random.default_rng(3)
    D = 6
    A = rng.standard_normal((D, D)); A = A @ A.T + 0.5 * np.eye(D)
    c = rng.standard_normal(D) * 0.7                      # some coordinates end on a bound, the others inside the box
    fg = lambda x: (float(-0.5 * (x - c) @ A @ (x - c) - 0.05 * np.sum((x - c) ** 4)), -A @ (x - c) - 0.2 * (x - c) ** 3)
    lb, ub = -np.ones(D), np.ones(D)
    r = lo.maximize(fg, np.zeros(D), lb, ub, maxeval=500, ftol_rel=1e-14)
    from scipy.optimize import minimize
    ref = minimize(lambda x: (-fg(x)[0], -fg(x)[1]), np.zeros(D), jac=True, method="L-BFGS-B", bounds=list(zip(lb, ub)), options=dict(ftol=1e-15, gtol=1e-12))
    assert np.all(r.x >= lb) and np.all(r.x <= ub) and r.status in (lo.FTOL, lo.XTOL)
    assert abs(r.f + ref.fun) <= 1e-8 * abs(ref.fun) and np.allclose(r.x, ref.x, atol=1e-5)
    r2 = lo.maximize(fg, np.zeros(D), lb, ub, maxeval=7)
    assert r2.status == lo.MAXEVAL and r2.evals == 7 and r2.f <= r.f + 1e-12
    r3 = lo.maximize(fg, np.zeros(D), lb, ub, maxeval=500, ftol_abs=1e-2)
    assert r3.status == lo.FTOL and r3.evals <= r.evals and r3.f <= r.f + 1e-12
    r4 = lo.maximize(fg, np.zeros(D), lb, ub, maxeval=500, xtol_abs=1e-3)
    assert r4.status in (lo.XTOL, lo.FTOL) and r4.evals <= r.evals
