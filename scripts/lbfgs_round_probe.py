"""Developer probe: the kernels of one L-BFGS lock-step round (16 starts) at N=1000 for an ncu launch list."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200bo as bo
from oracle import gp_oracle as orc
rng = np.random.default_rng(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
X = rng.random((2, N)) * 15.0 + np.array([[-5.0], [0.0]]); y = -np.array([orc.branin(X[0, i], X[1, i]) for i in range(N)])
m = bo.ElasticGPE(2, mean=bo.MeanConst(-10.0), kernel=bo.SEArd([1.0, 1.0], 4.0), logNoise=-2.0, capacity=3000)
m.append(X, y)
lb, ub = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
Xs = bo.ScaledLHSIterator(lb, ub, 16, np.random.default_rng(1)).data
r = m.acquire_lbfgs("EI", (float(y.max()),), Xs, lb, ub, maxeval=12)
print(r["evals"].max())
