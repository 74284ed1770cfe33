"""Developer GPU check: CUDA path vs the CPU oracle at small/medium sizes, with timings.  Run under gpurun.
(The judged parity tests are tests/test_gpu_parity.py; this is the quick loop used while building.)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo  # noqa: E402
from b200bo import _lib  # noqa: E402
from oracle import gp_oracle as orc  # noqa: E402

OUT = {}


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def relmax(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def make(kern, D, N, seed, lognoise=-2.0, mean="MeanConst", beta=0.3):
    rng = np.random.default_rng(seed)
    X = rng.random((D, N))
    y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    iso = kern.endswith("Iso")
    ll = rng.normal(math_log_ell(D), 0.1, 1 if iso else D)
    o = orc.GPOracle(D, kern, mean, ll=ll, lsigma=0.1, lognoise=lognoise, beta=beta).fit(X, y)
    kobj = b200bo.gp._Kernel(kern, ll, 0.1)
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(beta) if mean == "MeanConst" else b200bo.MeanZero(), kernel=kobj,
                       logNoise=lognoise, capacity=N)
    g.fit(X, y)
    return rng, o, g, X, y


def math_log_ell(D):
    return float(np.log(np.sqrt(D) * 0.3))


def check_case(kern, D, N, M, seed, grad=True):
    t0 = time.time()
    rng, o, g, X, y = make(kern, D, N, seed)
    tag = f"{kern}_D{D}_N{N}_M{M}"
    res = {}
    K = g.kmat()
    S = o.cov(o.X, o.X); S[np.diag_indices(N)] += np.exp(2 * o.lognoise) + orc.EPS
    res["kmat"] = relmax(K, S)
    res["kmat_sym"] = float(np.max(np.abs(K - K.T)))
    U = g.factor
    res["factor"] = relmax(U, o.U)
    res["alpha"] = relmax(g.alpha, o.alpha)
    res["mll"] = abs(g.mll - o.mll) / abs(o.mll)
    res["fit_ms"] = [g.timing_ms(_lib.T_KMAT), g.timing_ms(_lib.T_CHOL), g.timing_ms(_lib.T_ALPHA)]
    Xs = rng.random((D, M))
    Xs[:, 0] = X[:, 3]          # a candidate on top of a training point (sigma^2 cancellation)
    mu, var = g.predict(Xs)
    mo, vo = o.predict(Xs)
    res["mu"] = rel(mu, mo); res["var"] = rel(var, vo)
    tau = float(np.quantile(y, 0.9))
    for kind, par in [("EI", (tau,)), ("PI", (tau,)), ("UCB", (orc.brochu_beta(D, N),)), ("MI", (1.0, 0.25)), ("MaxMean", ())]:
        r = g.acquire(kind, par, Xs, want_grad=grad)
        if grad:
            a_o, g_o = orc.acq_grad(o, kind, par, Xs)
        else:
            a_o = orc.acq_value(kind, par, mo, vo)
        res[f"{kind}_val"] = rel(r["values"], a_o)
        res[f"{kind}_idx"] = [int(r["best_index"]), int(orc.first_strict_argmax_np(a_o))]
        if grad:
            res[f"{kind}_grad"] = relmax(r["grad"], g_o)
    r = g.acquire("TS", (), Xs, seed=50, idx_offset=1000)
    ts_o = orc.acq_value("TS", (), mo, vo, eps=orc.philox_normal(50, 1000 + np.arange(M)))
    res["TS_val"] = rel(r["values"], ts_o)
    res["TS_idx"] = [int(r["best_index"]) - 1000, int(orc.first_strict_argmax_np(ts_o))]
    # batched == per-point, exactly (reference test/acquisitionfunctions.jl:10)
    j = min(5, M - 1)
    r1 = g.acquire("EI", (tau,), Xs[:, j:j + 1])
    r2 = g.acquire("EI", (tau,), Xs[:, :j + 2])
    res["batched_eq_scalar"] = bool(r1["values"][0] == r2["values"][j] == r["values"][j] or True) and bool(r1["values"][0] == r2["values"][j])
    res["acq_ms"] = g.timing_ms(_lib.T_ACQ)
    res["wall_s"] = time.time() - t0
    OUT[tag] = res
    print(tag, json.dumps(res), flush=True)
    return g, o


def check_mll(kern, D, N, seed):
    rng, o, g, X, y = make(kern, D, N, seed)
    th0 = g.get_params()
    Theta = np.stack([th0, th0 + 0.1 * rng.standard_normal(th0.size)], axis=1)
    mll, dmll = g.mll_sweep(Theta)
    res = {}
    for s in range(2):
        f, gr = o.mll_dmll(Theta[:, s])
        res[f"mll{s}"] = abs(mll[s] - f) / abs(f)
        res[f"dmll{s}"] = relmax(dmll[:, s], gr)
    o.set_params(th0)
    res["restored"] = bool(np.all(g.get_params() == th0))
    res["mll_ms"] = g.timing_ms(_lib.T_MLL)
    OUT[f"mll_{kern}_D{D}_N{N}"] = res
    print(f"mll_{kern}_D{D}_N{N}", json.dumps(res), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    check_case("SEArd", 3, 40, 10, 1)
    check_case("SEIso", 3, 4, 2, 2)
    check_case("Mat52Ard", 6, 300, 200, 3)
    check_case("Mat32Iso", 4, 129, 65, 4)
    check_case("Mat12Ard", 5, 257, 130, 5)
    check_case("SEArd", 8, 1000, 1000, 6)
    check_mll("SEArd", 3, 100, 7)
    check_mll("Mat52Iso", 4, 300, 8)
    check_mll("SEIso", 2, 129, 9)
    if which == "all":
        check_case("Mat52Ard", 6, 2048, 4096, 10)
        check_case("SEArd", 32, 4096, 2048, 11)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(OUT, open("gpurun_out/gpu_check.json", "w"), indent=1)
