"""Developer probe: where one BO iteration spends its time at small N (Branin, the reference's own example size)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo as bo
from oracle import gp_oracle as orc
rng = np.random.default_rng(0)
for N in (50, 200, 1000):
    X = rng.random((2, N)) * 15.0 + np.array([[-5.0], [0.0]]); y = -np.array([orc.branin(X[0, i], X[1, i]) for i in range(N)])
    m = bo.ElasticGPE(2, mean=bo.MeanConst(-10.0), kernel=bo.SEArd([1.0, 1.0], 4.0), logNoise=-2.0, capacity=3000)
    m.append(X, y)
    lb, ub = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    tau = float(y.max())
    Xs = bo.ScaledLHSIterator(lb, ub, 4096, np.random.default_rng(1)).data
    def T(f, n=5):
        f(); t0 = time.perf_counter()
        for _ in range(n): r = f()
        return (time.perf_counter() - t0) / n * 1e3, r
    t_sweep, r = T(lambda: m.acquire("EI", (tau,), Xs, want_values=True))
    top = np.argsort(-r["values"])[:16]
    t_lb, r2 = T(lambda: m.acquire_lbfgs("EI", (tau,), Xs[:, top], lb, ub, maxeval=2000))
    t_app, _ = T(lambda: (m.append(rng.random((2, 1)) * 5, np.array([-3.0]))), n=3)
    t_fit, _ = T(lambda: bo.gp.update(m, None, np.array([])), n=3)
    print(f"N={N}: sweep of 4096 candidates {t_sweep:.2f} ms | L-BFGS refine of 16 starts {t_lb:.2f} ms (evals max {int(r2['evals'].max())}) | append 1 point {t_app:.2f} ms | refit {t_fit:.2f} ms", flush=True)
