"""Pinned parity, when the fixtures exist: tests/golden/ref_<case>.json are outputs of the REAL reference (BayesianOptimization.jl +
GaussianProcesses.jl) on the golden inputs, produced by oracle/make_ref_fixtures.jl on a machine with Julia.  They are absent from a
fresh checkout of this build (no Julia in the image): the checks below then SKIP with the reason "PARITY UNPINNED", and one always-on test
keeps the generator, its inputs and this loader consistent so that a maintainer can pin parity with one command."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import gp_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REFS = sorted(glob.glob(os.path.join(GOLD, "ref_*.json")))
REFS = [f for f in REFS if not f.endswith("ref_inputs.json")]
UNPINNED = "PARITY UNPINNED: no tests/golden/ref_<case>.json -- run `julia oracle/make_ref_fixtures.jl` where the reference is installed"
RTOL_POST, RTOL_ACQ, ATOL = 1e-6, 1e-5, 1e-12


def close(a, b, rtol, atol=ATOL):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return bool(np.all(np.abs(a - b) <= rtol * np.abs(b) + atol))


def test_generator_and_inputs_are_in_place():
    """always on: the Julia generator exists, names every quantity this loader reads, and its input file covers every golden case"""
    src = open(os.path.join(ROOT, "oracle", "make_ref_fixtures.jl")).read()
    for key in ("alpha", "mll", "Udiag", "mu", "var", "_values", "_best", "dmll", "versions", "joint_cov", "joint_mu", "direct_maxmean_f", "direct_lb"):
        assert key in src
    inp = json.load(open(os.path.join(GOLD, "ref_inputs.json")))
    names = {c["name"] for c in inp["cases"]}
    assert names == {os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "*.npz"))}
    for c in inp["cases"]:
        z = np.load(os.path.join(GOLD, c["name"] + ".npz"))
        assert np.array_equal(np.array(c["X"]), z["X"]) and np.array_equal(np.array(c["theta"]), z["theta"])
    if not REFS:
        print(UNPINNED)


def _case(path):
    ref = json.load(open(path))
    z = np.load(os.path.join(GOLD, ref["name"] + ".npz"))
    return ref, z


@pytest.mark.skipif(not REFS, reason=UNPINNED)
@pytest.mark.parametrize("path", REFS or [None])
def test_cpu_restatement_against_the_reference(path):
    ref, z = _case(path)
    th = z["theta"]; nm = 1 if str(z["mean"]) == "MeanConst" else 0
    o = orc.GPOracle(int(z["D"]), str(z["kernel"]), str(z["mean"]), ll=th[1 + nm:-1], lsigma=th[-1], lognoise=th[0], beta=th[1] if nm else 0.0).fit(z["X"], z["y"])
    assert close(o.alpha, ref["alpha"], 1e-8, 1e-10) and abs(o.mll - ref["mll"]) <= 1e-9 * abs(ref["mll"])
    mu, var = o.predict(z["Xs"])
    assert close(mu, ref["mu"], RTOL_POST) and close(var, ref["var"], RTOL_POST)
    for k in ("EI", "PI", "UCB", "MI", "MaxMean"):
        a = orc.acq_value(k, tuple(z[f"{k}_params"]), mu, var)
        assert close(a, ref[f"{k}_values"], RTOL_ACQ), k
        assert orc.first_strict_argmax_np(a) == ref[f"{k}_best"], k
    for key, t in (("", th), ("2", z["theta2"])):
        f, g = o.mll_dmll(t)
        assert abs(f - ref["mll" + key]) <= 1e-9 * abs(ref["mll" + key]) and close(g, ref["dmll" + key], 1e-6, 1e-9)
    if "joint_cov" in ref:                       # the covariance behind the joint draw of myrand(model, X::Matrix) (gp.jl:7)
        m = int(ref["joint_m"])
        o.set_params(th); o.fit(z["X"], z["y"])
        mj, Sj = o.posterior_cov(z["Xs"][:, :m])
        Sr = np.array(ref["joint_cov"], float)
        assert close(mj, ref["joint_mu"], RTOL_POST) and np.abs(Sj - Sr).max() <= 1e-6 * np.abs(np.diag(Sr)).max()


@pytest.mark.gpu
@pytest.mark.skipif(not REFS, reason=UNPINNED)
@pytest.mark.parametrize("path", REFS or [None])
def test_cuda_path_against_the_reference(path):
    import b200bo as bo
    ref, z = _case(path)
    th = z["theta"]; nm = 1 if str(z["mean"]) == "MeanConst" else 0
    g = bo.B200GPE(int(z["D"]), mean=bo.MeanConst(th[1]) if nm else bo.MeanZero(), kernel=bo.gp._Kernel(str(z["kernel"]), th[1 + nm:-1], th[-1]),
                   logNoise=th[0], capacity=z["y"].size)
    g.fit(z["X"], z["y"])
    assert close(g.alpha, ref["alpha"], 1e-8, 1e-10) and abs(g.mll - ref["mll"]) <= 1e-9 * abs(ref["mll"])
    mu, var = g.predict(z["Xs"])
    assert close(mu, ref["mu"], RTOL_POST) and close(var, ref["var"], RTOL_POST)
    for k in ("EI", "PI", "UCB", "MI", "MaxMean"):
        r = g.acquire(k, z[f"{k}_params"], z["Xs"])
        assert close(r["values"], ref[f"{k}_values"], RTOL_ACQ), k
        assert r["best_index"] == ref[f"{k}_best"], k                          # selected index bit-exact
    mll, dmll = g.mll_sweep(np.stack([th, z["theta2"]], axis=1))
    for j, key in enumerate(("", "2")):
        assert abs(mll[j] - ref["mll" + key]) <= 1e-9 * abs(ref["mll" + key]) and close(dmll[:, j], ref["dmll" + key], 1e-6, 1e-9)
    if "joint_cov" in ref:                       # b200bo_rand_joint == mu_ref + chol(Sigma_ref) eps for the library's own Philox normals
        from scipy.linalg import cholesky
        m = int(ref["joint_m"])
        Sr = np.array(ref["joint_cov"], float)
        r = g.rand_joint(z["Xs"][:, :m], seed=5)
        if r["tries"] == 0:
            want = np.array(ref["joint_mu"], float) + cholesky(Sr, lower=True) @ orc.philox_normal(5, np.arange(m))
            assert np.abs(r["sample"] - want).max() <= 1e-6 * np.sqrt(np.abs(np.diag(Sr)).max())
    if "direct_maxmean_f" in ref:                # NLopt's GN_DIRECT_L optimum of the posterior mean over the same box, same budget
        rd = g.acquire_direct("MaxMean", (), np.array(ref["direct_lb"], float), np.array(ref["direct_ub"], float), maxeval=2000)
        span = float(np.max(ref["mu"]) - np.min(ref["mu"]))
        assert rd["best_value"] >= ref["direct_maxmean_f"] - 5e-3 * span
