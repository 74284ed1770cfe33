"""Developer A/B: the MAP sweep (cfg4: N=4096, D=8, 64 settings) under the Cholesky schedules, by number of worker models in flight."""
import time, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo as bo
rng = np.random.default_rng(4)
X = rng.random((8, 4096)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(4096)
grid = [(a, l) for a in np.linspace(-3, 0, 8) for l in np.linspace(-1.5, 0.5, 8)]
Theta = np.stack([np.concatenate([[a, 0.0], np.full(8, l), [0.0]]) for a, l in grid], axis=1)
for sched, graph in ((0, 0), (1, 0), (1, 1)):
    for K in (1, 3, 6, 8):
        m = bo.B200GPE(8, mean=bo.MeanConst(0.0), kernel=bo.SEArd(np.zeros(8), 0.0), logNoise=-2.0, capacity=4096)
        m.fit(X, y)
        m.set_knob("sweep_workers", K); m.set_knob("chol_sched", sched); m.set_knob("chol_graph", graph)
        m.mll_sweep(Theta[:, :3 * K])
        t0 = time.perf_counter(); a, b = m.mll_sweep(Theta); t1 = time.perf_counter() - t0
        t0 = time.perf_counter(); a2, _ = m.mll_sweep(Theta, want_grad=False); t2 = time.perf_counter() - t0
        print(f"sched={sched} graph={graph} workers={K}: with gradient {t1:.3f} s, values only {t2:.3f} s", flush=True)
        del m
