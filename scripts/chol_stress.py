"""Developer check: repeated factorisations must be bitwise identical (cross-stream ordering bugs show up as run-to-run differences)."""
import sys, os, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo
for N, D in ((4096, 8), (2176, 5), (8192, 16)):
    rng = np.random.default_rng(N)
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.SEArd(np.full(D, np.log(np.sqrt(D) * 0.25)), 0.0), logNoise=-2.0, capacity=N)
    hs = set()
    for it in range(6 if N < 8192 else 3):
        g.fit(X, y)
        hs.add((hashlib.sha1(g.factor.tobytes()).hexdigest(), hashlib.sha1(g.alpha.tobytes()).hexdigest(), g.mll))
    print(N, "distinct results over refits:", len(hs), flush=True)
    assert len(hs) == 1
print("stress ok")
