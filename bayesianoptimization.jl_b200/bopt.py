"""BO driver surface of the reference (src/BayesianOptimization.jl): BOpt, boptimize!, optimize.

Host glue (O(1) work per iteration around the user's function); the hot path it calls -- setparams!, acquire_max,
update!, optimizemodel!, acquire_model_max -- runs on the B200 through the C ABI.
"""
from __future__ import annotations

import enum
import time

import numpy as np

from . import gp as _gp
from .acquisition import acquire_max, acquire_model_max, defaultoptions, nlopt_setup
from .acquisitionfunctions import ExpectedImprovement, setparams
from .gp import B200GPE, MAPGPOptimizer, MeanConst, Mat52Ard, optimizemodel, update
from .utils import DurationCounter, IterationCounter, ScaledSobolIterator, init, isdone, step


class Sense(enum.IntEnum):            # BayesianOptimization.jl:56
    Min = -1
    Max = 1


class Verbosity(enum.IntEnum):        # :57
    Silent = 0
    Timings = 1
    Progress = 2


Min, Max = Sense.Min, Sense.Max
Silent, Timings, Progress = Verbosity.Silent, Verbosity.Timings, Verbosity.Progress


class TimerOutput:
    """The four @mytimeit sections of the reference (utils.jl:1-7; BayesianOptimization.jl:169-170,185,194-200,210)."""

    def __init__(self):
        self.sections = {}

    def reset(self):
        self.sections = {}

    def add(self, name, dt):
        n, t = self.sections.get(name, (0, 0.0))
        self.sections[name] = (n + 1, t + dt)

    def __str__(self):
        return "\n".join(f"{k:40s} ncalls={n:6d} time={t:.4f}s" for k, (n, t) in self.sections.items())


class _timeit:
    def __init__(self, to, name):
        self.to, self.name = to, name

    def __enter__(self):
        self.t0 = time.perf_counter()

    def __exit__(self, *a):
        self.to.add(self.name, time.perf_counter() - self.t0)


class BOpt:
    """BOpt(func, model, acquisition, modeloptimizer, lowerbounds, upperbounds; kwargs...) (:59-136)."""

    def __init__(self, func, model, acquisition, modeloptimizer, lowerbounds, upperbounds, sense=Max, maxiterations=10 ** 4,
                 maxduration=float("inf"), acquisitionoptions=None, repetitions=1, verbosity=Progress,
                 initializer_iterations=None, initializer=None):
        now = time.time()
        lowerbounds = np.asarray(lowerbounds, float); upperbounds = np.asarray(upperbounds, float)
        if initializer_iterations is None:
            initializer_iterations = 5 * lowerbounds.size
        acquisitionoptions = {**defaultoptions(type(model), type(acquisition)), **(acquisitionoptions or {})}   # :105-106
        if lowerbounds.size != upperbounds.size:                                                                 # :112
            raise ValueError("length of lowerbounds does not match length of upperbounds")
        if initializer is None:
            initializer = ScaledSobolIterator(lowerbounds, upperbounds, initializer_iterations)
        if maxiterations < len(initializer):                                                                     # :107
            raise ValueError(f"maxiterations = {maxiterations} < length(initializer) = {len(initializer)}")
        if not maxiterations >= 0:
            raise ValueError("maxiterations < 0")
        if not maxduration >= 0:
            raise ValueError("maxduration < 0")
        if not np.all(lowerbounds <= upperbounds):                                                               # :114
            raise ValueError("lowerbounds are not pointwise less than or eqal to upperbounds, they were possibly "
                             "passed in the wrong order")
        y = model.y
        if y.size == 0:                                                                                          # :117-119
            current_optimum = -np.inf * int(sense)
            current_optimizer = np.zeros_like(lowerbounds)
        else:
            current_optimum = int(sense) * float(np.max(y))
            current_optimizer = np.array(model.x[:, int(np.argmax(y))])
        self.func, self.sense, self.model, self.acquisition = func, Sense(sense), model, acquisition
        self.acquisitionoptions, self.modeloptimizer = acquisitionoptions, modeloptimizer
        self.lowerbounds, self.upperbounds = lowerbounds, upperbounds
        self.observed_optimum, self.observed_optimizer = current_optimum, current_optimizer
        self.model_optimum, self.model_optimizer = current_optimum, current_optimizer.copy()
        self.iterations = IterationCounter(0, 0, maxiterations)
        self.duration = DurationCounter(now, maxduration, now, now + maxduration)
        self.opt = nlopt_setup(acquisition, model, lowerbounds, upperbounds, acquisitionoptions)                 # :134
        self.verbosity, self.initializer, self.repetitions = Verbosity(verbosity), initializer, repetitions
        self.timeroutput = TimerOutput()

    def __repr__(self):                                                                                          # :141-157
        s = f"Bayesian Optimization object\n\nmodel: B200GPE(D={self.model.D}, nobs={self.model.nobs})\n" \
            f"acquisition: {type(self.acquisition).__name__}\n"
        if self.iterations.i == 0:
            return s + "\nNo observation data."
        return s + (f"\nobserved optimum: {self.observed_optimum}\nobserved optimizer: {self.observed_optimizer}\n"
                    f"model optimum: {self.model_optimum}\nmodel optimizer: {self.model_optimizer}\n"
                    f"iterations: {self.iterations.i}/{self.iterations.N}\n"
                    f"duration: {self.duration.now - self.duration.starttime}/{self.duration.duration} s")


def maxduration(o: BOpt, d):
    o.duration.duration = d


def maxiterations(o: BOpt, N):
    o.iterations.N = N


def _evaluate_function(o: BOpt, x):                                                                              # :209-216
    with _timeit(o.timeroutput, "function evaluation"):
        y = int(o.sense) * o.func(x)
    if y > int(o.sense) * o.observed_optimum:
        o.observed_optimum = int(o.sense) * y
        o.observed_optimizer = x
    return y


def initialise_model(o: BOpt):                                                                                   # :159-172
    ys, xs = [], []
    for x in o.initializer:
        for _ in range(o.repetitions):
            ys.append(_evaluate_function(o, x))
            xs.append(x)
    o.iterations.i = o.iterations.c = len(ys) // o.repetitions
    with _timeit(o.timeroutput, "model update"):
        update(o.model, np.stack(xs, axis=1), np.array(ys))
    with _timeit(o.timeroutput, "model hyperparameter optimization"):
        optimizemodel(o.modeloptimizer, o.model)


def boptimize(o: BOpt):
    """boptimize!(o) (:176-207)."""
    init(o.duration)
    init(o.iterations)
    o.timeroutput.reset()
    if o.iterations.i == 0 and len(o.initializer) > 0:
        initialise_model(o)
    while not (isdone(o.iterations) or isdone(o.duration)):
        if o.verbosity >= Progress:
            print(f"{time.strftime('%FT%T')}\titeration: {o.iterations.i}\tcurrent optimum: {o.observed_optimum}")
        setparams(o.acquisition, o.model)                                                                        # :184
        with _timeit(o.timeroutput, "acquisition"):
            f, x = acquire_max(o.opt, o.lowerbounds, o.upperbounds, o.acquisitionoptions["restarts"])            # :185-187
        ys = []
        step(o.iterations)
        for _ in range(o.repetitions):
            ys.append(_evaluate_function(o, x))
        with _timeit(o.timeroutput, "model update"):
            update(o.model, np.stack([x] * o.repetitions, axis=1), np.array(ys))                                 # :194-196
        with _timeit(o.timeroutput, "model hyperparameter optimization"):
            optimizemodel(o.modeloptimizer, o.model)                                                             # :197-198
    with _timeit(o.timeroutput, "acquisition"):
        o.model_optimum, o.model_optimizer = acquire_model_max(o)                                                # :200
    o.duration.now = time.time()
    if o.verbosity >= Timings:
        print(o.timeroutput)
    return dict(observed_optimum=o.observed_optimum, observed_optimizer=o.observed_optimizer,
                model_optimum=int(o.sense) * o.model_optimum, model_optimizer=o.model_optimizer)


_ARGS_KEYS = ("model", "acquisition", "modeloptimizer")
_KWARGS_KEYS = ("sense", "maxiterations", "maxduration", "acquisitionoptions", "repetitions", "verbosity",
                "initializer_iterations", "initializer")


def merge_with_defaults(f, lowerbounds, upperbounds, optkwargs: dict):
    """merge_with_defaults (:238-289): same argument order and ArgumentErrors (ValueError here)."""
    if not set(optkwargs) <= set(_ARGS_KEYS) | set(_KWARGS_KEYS):
        raise ValueError("use of unsupported keyword arguments")
    if len(lowerbounds) != len(upperbounds):
        raise ValueError("length of lowerbounds does not match length of upperbounds")
    D = len(lowerbounds)
    params = dict(optkwargs)
    if "model" not in params:                                                                                   # :259-264
        params["model"] = B200GPE(D, mean=MeanConst(0.0), kernel=Mat52Ard(np.zeros(D), 0.0), logNoise=-2.0, capacity=3000)
    params.setdefault("acquisition", ExpectedImprovement())
    if "modeloptimizer" not in params:                                                                          # :266-272
        params["modeloptimizer"] = MAPGPOptimizer(every=20, noisebounds=[-4, 3],
                                                  kernbounds=[[-3.0] * D + [-3.0], [4.0] * D + [3.0]], maxeval=100)
    params.setdefault("maxiterations", 10 ** 3)
    args = (f, *[params[k] for k in _ARGS_KEYS], lowerbounds, upperbounds)
    kwargs = {k: v for k, v in params.items() if k in _KWARGS_KEYS}
    return args, kwargs


def optimize(f, lowerbounds, upperbounds, **optkwargs):
    """optimize(f, lowerbounds, upperbounds; kwargs...) (:230-234)."""
    args, kwargs = merge_with_defaults(f, lowerbounds, upperbounds, optkwargs)
    return boptimize(BOpt(*args, **kwargs))
