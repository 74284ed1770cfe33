"""cost of b200bo_append for 1 vs 16 points right after an acquisition (W = L^-1 available): python scripts/append_bench.py [N]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo as bo
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
D = 6
rng = np.random.default_rng(1)
X = rng.random((D, N + 400)); y = np.sin(3 * X.sum(0))
Xs = rng.random((D, 2048))
for m in (1, 16, 1, 16):
    g = bo.B200GPE(D, mean=bo.MeanConst(0.0), kernel=bo.Mat52Ard(np.zeros(D), 0.0), logNoise=-2.0, capacity=N + 512)
    g.fit(X[:, :N], y[:N])
    ts = []
    n = N
    for rep in range(6):
        g.acquire("EI", (0.5,), Xs)
        t0 = time.perf_counter(); g.append(X[:, n:n + m], y[n:n + m]); ts.append(time.perf_counter() - t0); n += m
    print(f"N={N}: append of {m:2d} point(s): median {np.median(ts[1:]) * 1e3:.3f} ms")
