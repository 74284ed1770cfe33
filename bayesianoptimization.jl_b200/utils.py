"""Host utilities of the reference (src/utils.jl): counters and the candidate / initialiser generators."""
from __future__ import annotations

import time

import numpy as np


class IterationCounter:                      # utils.jl:9-23
    def __init__(self, c: int, i: int, N: int):
        self.c, self.i, self.N = c, i, N


class DurationCounter:                       # utils.jl:25-42
    def __init__(self, starttime, duration, now, endtime):
        self.starttime, self.duration, self.now, self.endtime = starttime, duration, now, endtime


def isdone(s) -> bool:
    if isinstance(s, IterationCounter):
        return s.c == s.N
    s.now = time.time()
    return s.now >= s.endtime


def step(s: IterationCounter):
    s.c += 1
    s.i += 1


def init(s):
    if isinstance(s, IterationCounter):
        s.c = 0
    else:
        s.starttime = time.time()
        s.endtime = s.starttime + s.duration


def latin_hypercube_sampling(mins, maxs, n: int, rng: np.random.Generator = None) -> np.ndarray:
    """utils.jl:101-120: per dimension n jittered strata, shuffled.  Returns D x n."""
    mins = np.asarray(mins, float); maxs = np.asarray(maxs, float)
    if mins.size != maxs.size:
        raise ValueError("mins and maxs should have the same length")          # DimensionMismatch
    if not np.all(mins <= maxs):
        raise ValueError("mins[i] should not exceed maxs[i]")                  # ArgumentError
    rng = np.random.default_rng() if rng is None else rng
    out = np.zeros((mins.size, n), order="F")
    for i in range(mins.size):
        dimstep = (maxs[i] - mins[i]) / n
        cube = mins[i] + dimstep * (np.arange(n) + rng.random(n))
        rng.shuffle(cube)
        out[i, :] = cube
    return out


class ScaledLHSIterator:                     # utils.jl:96-98 (a ColumnIterator over the LHS matrix)
    def __init__(self, lowerbounds, upperbounds, N: int, rng=None):
        self.data = latin_hypercube_sampling(lowerbounds, upperbounds, N, rng)

    def __len__(self):
        return self.data.shape[1]

    def __iter__(self):
        return (self.data[:, j] for j in range(self.data.shape[1]))


class ScaledSobolIterator:                   # utils.jl:64-87
    """N points of the (unscrambled, Joe-Kuo) Sobol sequence scaled to the box.  As in the reference, the constructor skips the
    start of the sequence with Sobol.jl's `skip(seq, N)` (utils.jl:80), which -- `exact=false` -- advances by the largest power of two
    2^floor(log2(N+1)) <= N+1, not by N; Sobol.jl never emits the origin, hence the extra 1.  Each pass over the iterator draws the
    NEXT N points of the shared sequence (utils.jl:84-87 calls `next!` on `it.seq`), it does not replay the first pass."""

    def __init__(self, lowerbounds, upperbounds, N: int):
        from scipy.stats import qmc
        self.lowerbounds = np.asarray(lowerbounds, float)
        self.upperbounds = np.asarray(upperbounds, float)
        self.N = int(N)
        self._seq = qmc.Sobol(self.lowerbounds.size, scramble=False)
        self._seq.fast_forward(1)                       # Sobol.jl never emits the origin
        if N > 0:
            self._seq.fast_forward(1 << int(np.floor(np.log2(N + 1))))

    def __len__(self):
        return self.N

    def __iter__(self):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            u = self._seq.random(self.N) if self.N > 0 else np.zeros((0, self.lowerbounds.size))
        pts = self.lowerbounds + u * (self.upperbounds - self.lowerbounds)
        return (pts[j].copy() for j in range(self.N))
