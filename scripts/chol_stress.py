"""Developer check: repeated factorisations must be bitwise identical (cross-stream ordering bugs show up as run-to-run differences)."""
import sys, os, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo
for N, D in ((4096, 8), (2176, 5), (8192, 16)):
    rng = np.random.default_rng(N)
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.SEArd(np.full(D, np.log(np.sqrt(D) * 0.25)), 0.0), logNoise=-2.0, capacity=N)
    g.set_knob("chol_graph", 2)
    hs = set()
    for it in range(6 if N < 8192 else 3):
        g.fit(X, y)
        hs.add((hashlib.sha1(g.factor.tobytes()).hexdigest(), hashlib.sha1(g.alpha.tobytes()).hexdigest(), g.mll))
    print(N, "distinct results over refits:", len(hs), flush=True)
    assert len(hs) == 1
print("stress ok")

# graph cache under changing shapes: every result must equal the eager in-order factorisation of the same data
rng = np.random.default_rng(7)
D = 4
Xall = rng.random((D, 2600)); yall = np.sin(3 * Xall.sum(0)) + 0.1 * rng.standard_normal(2600)
g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.2), kernel=b200bo.Mat52Ard(np.full(D, -0.5), 0.1), logNoise=-2.0, capacity=2600)
ref = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.2), kernel=b200bo.Mat52Ard(np.full(D, -0.5), 0.1), logNoise=-2.0, capacity=2600)
ref.set_knob("chol_sched", 0)
g.set_knob("chol_graph", 2)
seq = [1000, 1000, 1000, 1001, 1200, 1200, 1000, 1000, 2600, 2600, 2600, 1200, 1000, 1000, 1023, 1024, 1025, 1025, 1025]
for n in seq:
    g.fit(Xall[:, :n], yall[:n]); ref.fit(Xall[:, :n], yall[:n])
    da = np.abs(g.alpha - ref.alpha).max() / np.abs(ref.alpha).max()
    assert da < 1e-9 and abs(g.mll - ref.mll) < 1e-9 * abs(ref.mll), (n, da)
print("graph cache under changing shapes ok:", len(seq), "fits")
