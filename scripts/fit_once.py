import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo
from b200bo import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.default_rng(0)
X = rng.random((D, N)); y = np.sin(3 * X.sum(0)) + 0.1 * rng.standard_normal(N)
g = b200bo.B200GPE(D, mean=b200bo.MeanConst(0.0), kernel=b200bo.SEArd(np.full(D, np.log(np.sqrt(D) * 0.25)), 0.0), logNoise=-2.0, capacity=N)
g.set_knob("chol_graph", 2)      # capture the factorisation at the second fit (default: the sixth consecutive one of a shape)
for _ in range(4):
    g.fit(X, y)
    print(json.dumps(dict(N=N, kmat_ms=g.timing_ms(_lib.T_KMAT), chol_ms=g.timing_ms(_lib.T_CHOL), syrk_ms=g.timing_ms(_lib.T_SYRK), alpha_ms=g.timing_ms(_lib.T_ALPHA))))
