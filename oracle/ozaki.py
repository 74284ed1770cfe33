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


# ---------------------------------------------------------------------------------------------------------------------------------
# K6 on tcgen05 (csrc/acq_i8.cu): the acquisition step as an error-free int8-slice GEMM against the explicit inverse factor,
#     v = U^-T k*  (EXT GP.predict_f, reached from src/models/gp.jl:2-5,8)  =  W k*,   W = L^-1,   sigma^2 = max(sf2 - |v|^2, 0).
# Digits are taken from a 55-bit integer here (slice_rows_kernel / kstar_slice_kernel): q = rint(x 2^(54-e)), then balanced radix-256
# digits through a bias add and byte extraction.
# ---------------------------------------------------------------------------------------------------------------------------------
BIAS = 0x0000808080808080


def int_digits(q):
    """q int64 (|q| <= 2^54) -> digits int64 [S][...]: q = d0 2^48 + sum_{p>=1} d_p 2^(8(6-p)), d_p in [-128,127], d0 in [-64,64]
    (i8_digits + gather_byte of csrc/umma.cuh / acq_i8.cu: byte k of (q + BIAS) ^ BIAS is slice 6 - k as an int8)."""
    q = np.asarray(q, np.int64)
    b = ((q + BIAS).astype(np.uint64)) ^ np.uint64(BIAS)
    out = np.empty((S,) + q.shape, np.int64)
    for s in range(S):
        byte = ((b >> np.uint64(8 * (6 - s))) & np.uint64(0xFF)).astype(np.int64)
        out[s] = np.where(byte >= 128, byte - 256, byte)
    return out


def slice_rows_int(X, tri_block=None):
    """X [rows][K] -> (digits [S][rows][K], scale [rows] = 2^(e-6)); per-row exponent from the row maximum.  tri_block = 128: row r only
    has columns [0, 128 (r // 128 + 1)) (W = L^-1 is lower triangular; the rest is never read)."""
    X = np.asarray(X, float)
    if tri_block:
        X = X.copy()
        for r in range(X.shape[0]):
            X[r, tri_block * (r // tri_block + 1):] = 0.0
    m = np.max(np.abs(X), axis=1)
    e = np.where(m > 0, np.frexp(np.where(m > 0, m, 1.0))[1], 0)
    q = np.rint(X * np.ldexp(1.0, 54 - e)[:, None]).astype(np.int64)
    return int_digits(q), np.ldexp(1.0, e - 6)


def kstar_slices(Ks, sf2):
    """k* [cands][N] (values in [0, sf2]) -> digits with the FIXED scale S = 2^e2 > sf2 (frexp), and sBk = S 2^-6."""
    e2 = int(np.frexp(sf2)[1])
    q = np.rint(np.asarray(Ks, float) * np.ldexp(1.0, 54 - e2)).astype(np.int64)
    return int_digits(q), float(np.ldexp(1.0, e2 - 6))


def sliced_gemm(Da, Db):
    """exact anti-diagonal accumulators of A B^T from digits [S][m][K], [S][n][K] and their FP64 recombination (epilogue of
    acq_i8_gemm_kernel): two 64-bit integer groups, the small one first."""
    acc = [np.zeros((Da.shape[1], Db.shape[1]), np.int64) for _ in range(S)]
    for p in range(S):
        for q in range(S - p):
            acc[p + q] += Da[p] @ Db[q].T
    assert all(np.abs(a).max() < 2 ** 31 for a in acc)          # the int32 TMEM accumulators
    H0 = ((acc[0] * 256 + acc[1]) * 256 + acc[2]) * 256 + acc[3]
    H1 = (acc[4] * 256 + acc[5]) * 256 + acc[6]
    return H0.astype(float) * 2.0 ** -24 + H1.astype(float) * 2.0 ** -48


def posterior_var_i8(L, Ks, sf2):
    """sigma^2 for candidates from the lower factor L [N][N] (N a multiple of 128 here) and k* [cands][N] as the tcgen05 path takes
    it: W = L^-1, row slices, the exact products per 64-row tile over the k-blocks up to the diagonal, |v|^2 summed per 32 rows and
    then over the partials in ascending order."""
    import scipy.linalg as sl
    N = L.shape[0]
    assert N % 128 == 0
    W = sl.solve_triangular(L, np.eye(N), lower=True)
    Dw, we = slice_rows_int(W, tri_block=128)
    Dk, sBk = kstar_slices(Ks, sf2)
    ss = np.zeros(Ks.shape[0])
    for it in range(N // 64):
        kmax = 128 * (it // 2 + 1)
        rows = slice(64 * it, 64 * it + 64)
        acc = sliced_gemm(Dk[:, :, :kmax], Dw[:, rows, :kmax])                  # [cands][64]
        v = acc * (sBk * we[rows])[None, :]
        for half in range(2):
            part = np.zeros(Ks.shape[0])
            for c in range(32 * half, 32 * half + 32):
                part = part + v[:, c] * v[:, c]                                  # fma chain over the 32 columns of a warp
            ss = ss + part
    return np.maximum(sf2 - ss, 0.0)
