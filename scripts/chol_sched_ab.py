"""Developer A/B: the look-ahead schedule of the blocked Cholesky against the in-order one (same kernels): timing, agreement, bitwise repeatability."""
import sys, os, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo
from b200bo import _lib
cfgs = [(2048, 8), (4096, 8), (2176, 5), (8192, 16), (1300, 6), (640, 3), (128, 2), (257, 2)]
for N, D in cfgs:
    rng = np.random.default_rng(N)
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.SEArd(np.full(D, np.log(np.sqrt(D) * 0.25)), 0.0), logNoise=-2.0, capacity=N)
    res = {}
    for sched in (0, 2, 1):
        g.set_knob("chol_sched", sched); g.set_knob("chol_graph", 2)      # capture at the second fit of a shape (default: the sixth)
        hs, ts = set(), []
        for it in range(5 if N < 8192 else 3):
            g.fit(X, y)
            ts.append(g.timing_ms(_lib.T_CHOL))
            F = g.factor
            hs.add((hashlib.sha1(F.tobytes()).hexdigest(), hashlib.sha1(g.alpha.tobytes()).hexdigest(), g.mll))
        res[sched] = (F, g.alpha.copy(), g.mll, min(ts), len(hs))
    F0, a0, m0, t0, n0 = res[0]; F1, a1, m1, t1, n1 = res[1]
    print(f"N={N} D={D}: chol in-order {t0:.3f} ms, look-ahead/tile heads {res[2][3]:.3f} ms, look-ahead/fused head {t1:.3f} ms | distinct results {n0}/{res[2][4]}/{n1} | max|dU| {np.abs(F0 - F1).max():.2e} "
          f"max|dalpha| {np.abs(a0 - a1).max():.2e} (|alpha| {np.abs(a0).max():.2e}) dmll {abs(m0 - m1):.2e}", flush=True)
    assert n0 == 1 and n1 == 1
    assert np.abs(F0 - F1).max() < 1e-9 and abs(m0 - m1) < 1e-7 * abs(m0)
print("sched A/B ok")
