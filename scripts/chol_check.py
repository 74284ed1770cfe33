import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo
from oracle import gp_oracle as orc
for N in [int(a) for a in sys.argv[1:]]:
    D = 8
    rng = np.random.default_rng(0)
    X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
    ll = np.full(D, np.log(np.sqrt(D) * 0.25))
    o = orc.GPOracle(D, "SEArd", "MeanConst", ll=ll, lsigma=0.0, lognoise=-2.0, beta=0.0).fit(X, y)
    g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.SEArd(ll, 0.0), logNoise=-2.0, capacity=N)
    try:
        g.fit(X, y)
        U = g.factor
        err = np.abs(U - o.U)
        blk = err.reshape(N // 128, 128, N // 128, 128).max(axis=(1, 3)) if N % 128 == 0 else None
        print(N, "factor relerr", err.max() / np.abs(o.U).max(), "jitter", g.jitter_tries)
        if blk is not None and err.max() > 1e-9:
            bad = np.argwhere(blk > 1e-9)
            print(" first bad blocks (row=colblock of U -> [col, row] of L):", bad[:10].tolist())
    except Exception as e:
        print(N, "FAILED", e)
