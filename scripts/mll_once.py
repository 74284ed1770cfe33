"""Developer probe: ONE MAP objective + gradient evaluation at N=4096, D=8 on the model's own buffers (for ncu launch lists)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo as bo
rng = np.random.default_rng(4)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X = rng.random((8, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
m = bo.B200GPE(8, mean=bo.MeanConst(0.0), kernel=bo.SEArd(np.zeros(8), 0.0), logNoise=-2.0, capacity=N)
m.fit(X, y)
m.set_knob("sweep_workers", 0)
th = np.concatenate([[-2.0, 0.0], np.full(8, -0.5), [0.0]]).reshape(-1, 1)
for _ in range(3):
    t0 = time.perf_counter(); a, g = m.mll_sweep(th); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); a2, _ = m.mll_sweep(th, want_grad=False); t2 = time.perf_counter() - t0
    print(f"one setting: with gradient {t1 * 1e3:.2f} ms, value only {t2 * 1e3:.2f} ms")
