"""Developer perf probe: fit + acquisition timings at the BASELINE configs (CUDA events inside the library)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo
from b200bo import _lib

def run(kern, D, N, M, acq, par, grad, reps=3):
    rng = np.random.default_rng(0)
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    ll = np.full(1 if kern.endswith("Iso") else D, np.log(np.sqrt(D) * 0.25))
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.gp._Kernel(kern, ll, 0.0), logNoise=-2.0, capacity=N)
    t0 = time.time(); g.fit(X, y); t_fit = time.time() - t0
    g.fit(X, y)
    fit = [g.timing_ms(_lib.T_KMAT), g.timing_ms(_lib.T_CHOL), g.timing_ms(_lib.T_ALPHA)]
    Xs = rng.random((D, M))
    ms = []
    for _ in range(reps):
        t0 = time.time()
        r = g.acquire(acq, par, Xs, want_grad=grad)
        wall = time.time() - t0
        ms.append((g.timing_ms(_lib.T_ACQ), wall * 1e3))
    k = min(m[0] for m in ms)
    flops = M * float(N) ** 2 * (2 if grad else 1)
    out = dict(kern=kern, D=D, N=N, M=M, acq=acq, grad=grad, fit_ms=fit, fit_wall_first_s=t_fit, acq_ms=ms,
               cand_per_s=M / (k * 1e-3), tflops=flops / (k * 1e-3) * 1e-12, jitter=g.jitter_tries, best=r["best_index"])
    print(json.dumps(out), flush=True)
    return out

def fit_only(D=8, N=4096):
    rng = np.random.default_rng(0)
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.SEArd(np.full(D, np.log(np.sqrt(D) * 0.25)), 0.0), logNoise=-2.0, capacity=N)
    g.fit(X, y); g.fit(X, y)
    print(json.dumps(dict(N=N, D=D, kmat_ms=g.timing_ms(_lib.T_KMAT), chol_ms=g.timing_ms(_lib.T_CHOL), syrk_ms=g.timing_ms(_lib.T_SYRK),
                          alpha_ms=g.timing_ms(_lib.T_ALPHA))), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fit":
        fit_only(); fit_only(32, 4096); fit_only(16, 8192); sys.exit(0)
    res = []
    res.append(run("Mat52Ard", 6, 2048, 65536, "UCB", (9.12,), False))
    res.append(run("SEArd", 8, 2048, 65536, "EI", (1.0,), False))
    res.append(run("SEArd", 8, 4096, 32768, "EI", (1.0,), False))
    res.append(run("SEArd", 32, 4096, 32768, "EI", (1.0,), True, reps=2))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gpu_perf.json", "w"), indent=1)
