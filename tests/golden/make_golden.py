"""Generates tests/golden/*.npz from the CPU restatement (oracle/gp_oracle.py), seeded.

The reference's own tests hold no numeric vectors for this path and Julia cannot run here, so
these fixtures pin the ORACLE (regression) and give the -m gpu tests fixed inputs/outputs that travel to the GPU box.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gp_oracle as orc  # noqa: E402

CASES = [  # name, kernel, mean, D, N, M, seed, lognoise
    ("branin_cfg1", "SEArd", "MeanConst", 2, 50, 40, 101, -2.0),
    ("hartmann_small", "Mat52Ard", "MeanConst", 6, 200, 64, 102, -2.0),
    ("seiso_zero_mean", "SEIso", "MeanZero", 3, 4, 2, 103, -2.0),
    ("mat32iso_ragged", "Mat32Iso", "MeanConst", 4, 129, 65, 104, -1.0),
    ("mat12ard", "Mat12Ard", "MeanZero", 5, 130, 33, 105, -1.5),
    ("seard_d16", "SEArd", "MeanConst", 16, 256, 70, 106, -2.0),
]


def build(name, kern, mean, D, N, M, seed, lognoise):
    rng = np.random.default_rng(seed)
    if name == "branin_cfg1":                          # BASELINE config 1 (test/branin.jl:18-38 plumbing)
        lb, ub = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
        X = orc.latin_hypercube_sampling(lb, ub, N, rng)
        y = -orc.branin(X[0], X[1])
        ll, lsigma, beta = np.array([0.0, 0.0]), 5.0, -10.0
        Xs = orc.latin_hypercube_sampling(lb, ub, M, rng)
    elif name == "hartmann_small":
        X = rng.random((D, N)); y = -orc.hartmann6(X)
        ll, lsigma, beta = np.zeros(D), 0.0, 0.0
        Xs = orc.latin_hypercube_sampling(np.zeros(D), np.ones(D), M, rng)
    else:
        X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
        ll = rng.normal(np.log(np.sqrt(D) * 0.3), 0.1, 1 if kern.endswith("Iso") else D)
        lsigma, beta = 0.1, (0.3 if mean == "MeanConst" else 0.0)
        Xs = rng.random((D, M))
    Xs[:, 0] = X[:, min(3, N - 1)]                      # one candidate on a training point
    gp = orc.GPOracle(D, kern, mean, ll=ll, lsigma=lsigma, lognoise=lognoise, beta=beta).fit(X, y)
    out = dict(kernel=kern, mean=mean, D=D, X=np.asfortranarray(X), y=y, theta=gp.get_params(), Xs=np.asfortranarray(Xs),
               alpha=gp.alpha, mll=gp.mll, Udiag=np.diag(gp.U).copy(), Ucol_last=gp.U[:, -1].copy())
    mu, var = gp.predict(Xs)
    out["mu"], out["var"] = mu, var
    tau = float(np.quantile(y, 0.9))
    acqs = {"EI": (tau,), "PI": (tau,), "UCB": (orc.brochu_beta(D, N),), "MI": (1.0, 0.25), "MaxMean": ()}
    for k, p in acqs.items():
        a, g = orc.acq_grad(gp, k, p, Xs)
        out[f"{k}_params"] = np.array(p, float)
        out[f"{k}_values"], out[f"{k}_grad"] = a, np.asfortranarray(g)
        out[f"{k}_best"] = orc.first_strict_argmax_np(a)
    out["TS_seed"], out["TS_offset"] = 50, 1000
    out["TS_values"] = orc.acq_value("TS", (), mu, var, eps=orc.philox_normal(50, 1000 + np.arange(M)))
    th = gp.get_params()
    f, g = gp.mll_dmll(th)
    out["dmll"] = g
    th2 = th + 0.05 * rng.standard_normal(th.size)
    f2, g2 = gp.mll_dmll(th2)
    out["theta2"], out["mll2"], out["dmll2"] = th2, f2, g2
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "N", N, "M", M, "mll", gp.mll)


if __name__ == "__main__":
    for c in CASES:
        build(*c)
