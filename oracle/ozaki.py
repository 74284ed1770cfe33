"""TEST INFRASTRUCTURE ONLY: NumPy restatement of the error-free int8 slicing used by the tcgen05 trailing update
(bayesianoptimization.jl_b200/csrc/syrk_i8.cu).  It is not part of the reference (jbrea/BayesianOptimization.jl has no such step: the
reference reaches LAPACK dpotrf through GaussianProcesses.jl, SURVEY App. A "Fit"); it pins the arithmetic of OUR replacement of the
K = 512 trailing update  A_ij -= L_i,P L_j,P^T  so that the kernel's digit rule, scales and truncation are checked on the CPU.

    x_k = 2^(e-6) sum_{p<7} d_p[k] 2^(-8p),   d_0 in [-64, 64], d_p in [-128, 127]   (exact for |x| < 2^e, 54 bits)
    x.y ~ 2^(ex+ey-12) sum_{p+q<=6} 2^(-8(p+q)) <d_p, d'_q>                           (28 exact integer dot products)
"""
import numpy as np

S = 7
C = 128.0 / 255.0


def slice_rows(X, e=None):
    """X [rows][K] float64 -> (digits int64 [S][rows][K], scale [rows] = 2^(e-6)) with the kernel's digit rule.
    `e` (scalar or per row) fixes the exponent instead of taking it from the row maximum (|x| <= 2^e must hold)."""
    X = np.asarray(X, float)
    m = np.max(np.abs(X), axis=1)
    if e is None:
        e = np.where(m > 0, np.frexp(m)[1], 0)                # |x| < 2^e  (frexp: m = f 2^e, f in [0.5, 1))
    else:
        e = np.broadcast_to(np.asarray(e, int), m.shape)
        assert np.all(m <= np.ldexp(1.0, e))
    t = X * np.ldexp(1.0, 6 - e)[:, None]
    D = np.empty((S,) + X.shape, np.int64)
    for s in range(S):
        d = np.clip(np.floor(t + C), -128.0, 127.0)
        D[s] = d.astype(np.int64)
        t = (t - d) * 256.0
    return D, np.ldexp(1.0, e - 6)


def reconstruct(D, scale):
    acc = np.zeros(D.shape[1:], float)
    for s in range(S - 1, -1, -1):
        acc += D[s] * 2.0 ** (-8 * s)
    return acc * scale[:, None]


def product(A, B, ea=None, eb=None):
    """A [m][K], B [n][K] -> A B^T as the kernel computes it: anti-diagonal int accumulators, two 64-bit groups, row scales."""
    Da, sa = slice_rows(A, ea)
    Db, sb = slice_rows(B, eb)
    acc = [np.zeros((A.shape[0], B.shape[0]), np.int64) for _ in range(S)]
    for p in range(S):
        for q in range(S - p):
            acc[p + q] += Da[p] @ Db[q].T
    assert all(np.abs(a).max() < 2 ** 31 for a in acc)          # fits the int32 TMEM accumulators
    H0 = ((acc[0] * 256 + acc[1]) * 256 + acc[2]) * 256 + acc[3]
    H1 = (acc[4] * 256 + acc[5]) * 256 + acc[6]
    val = H1.astype(float) * 2.0 ** -48
    val = H0.astype(float) * 2.0 ** -24 + val
    return val * sa[:, None] * sb[None, :]
