"""Timings of the BASELINE configs 3-5 on one B200 (library CUDA-event timers).  Output: gpurun_out/configs.json"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo
from b200bo import _lib
from oracle import gp_oracle as orc

out = {}
def model(D, N, kern, seed):
    rng = np.random.default_rng(seed)
    X = rng.random((D, N))
    c = rng.random((D, 8))
    y = sum(np.exp(-0.5 * np.sum((X - c[:, k:k + 1]) ** 2, axis=0) / 0.15) for k in range(8)) + np.exp(-2.0) * rng.standard_normal(N)
    ll = np.full(D, np.log(np.sqrt(D) * 0.25))
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.gp._Kernel(kern, ll, 0.0), logNoise=-2.0, capacity=N)
    g.fit(X, y); g.fit(X, y)
    return g, X, y, rng

which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
if "cfg3" in which:   # D=32, N=4096, EI + gradient, M=262144
    D, N, M = 32, 4096, 262144
    g, X, y, rng = model(D, N, "SEArd", 3)
    Xs = orc.latin_hypercube_sampling(np.zeros(D), np.ones(D), M, np.random.default_rng(30))
    res = []
    for grad in (False, True):
        for _ in range(2):
            t0 = time.time(); r = g.acquire("EI", (float(y.max()),), Xs, want_grad=grad, want_values=True); wall = time.time() - t0
        ms = g.timing_ms(_lib.T_ACQ)
        fl = M * float(N) ** 2 * (2 if grad else 1)
        res.append(dict(grad=grad, kernel_ms=ms, wall_ms=wall * 1e3, cand_per_s=M / (ms * 1e-3), tflops=fl / (ms * 1e-3) * 1e-12, best=r["best_index"]))
    out["cfg3"] = dict(D=D, N=N, M=M, fit_ms=[g.timing_ms(k) for k in (_lib.T_KMAT, _lib.T_CHOL, _lib.T_SYRK, _lib.T_ALPHA)], runs=res)
    print(json.dumps(out["cfg3"]), flush=True)
if "cfg4" in which:   # MAP sweep: N=4096, D=8, 64 settings
    D, N = 8, 4096
    g, X, y, rng = model(D, N, "SEArd", 4)
    th0 = g.get_params()
    Theta = np.tile(th0[:, None], (1, 64))
    ln = np.linspace(-3, 0, 8); lls = np.linspace(-1.5, 0.5, 8)
    k = 0
    for a in ln:
        for b in lls:
            Theta[0, k] = a; Theta[2:2 + D, k] = b; k += 1
    t0 = time.time(); mll, dmll = g.mll_sweep(Theta); wall = time.time() - t0
    out["cfg4"] = dict(D=D, N=N, S=64, sweep_ms=g.timing_ms(_lib.T_MLL), wall_ms=wall * 1e3, per_setting_ms=g.timing_ms(_lib.T_MLL) / 64,
                       mll_range=[float(mll.min()), float(mll.max())], finite=bool(np.all(np.isfinite(dmll))))
    t0 = time.time(); mll2, _ = g.mll_sweep(Theta, want_grad=False); wall = time.time() - t0
    out["cfg4"]["value_only_ms"] = g.timing_ms(_lib.T_MLL)
    print(json.dumps(out["cfg4"]), flush=True)
if "cfg5" in which:   # D=16, N=8192, TS, M=1048576/8 per GPU
    D, N, M = 16, 8192, 131072
    g, X, y, rng = model(D, N, "SEArd", 5)
    Xs = orc.latin_hypercube_sampling(np.zeros(D), np.ones(D), M, np.random.default_rng(51))
    for _ in range(2):
        t0 = time.time(); r = g.acquire("TS", (), Xs, seed=50, want_values=False); wall = time.time() - t0
    ms = g.timing_ms(_lib.T_ACQ)
    out["cfg5"] = dict(D=D, N=N, M_per_gpu=M, fit_ms=[g.timing_ms(k) for k in (_lib.T_KMAT, _lib.T_CHOL, _lib.T_SYRK, _lib.T_ALPHA)], kernel_ms=ms,
                       wall_ms=wall * 1e3, cand_per_s=M / (ms * 1e-3), tflops=M * float(N) ** 2 / (ms * 1e-3) * 1e-12, best=r["best_index"])
    print(json.dumps(out["cfg5"]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
