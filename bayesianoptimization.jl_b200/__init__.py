"""b200bo -- B200-native GP-posterior + acquisition-search path behind the BayesianOptimization.jl surface.

The directory name (`bayesianoptimization.jl_b200`) is not a valid Python identifier; import it through the
repo-root shim:  `import b200bo`.
"""
from . import _lib                                   # raises if libb200bo.so is missing (no fallback)
from .gp import (B200GPE, MeanZero, MeanConst, SEIso, SEArd, Mat12Iso, Mat12Ard, Mat32Iso, Mat32Ard, Mat52Iso,
                 Mat52Ard, mean_var, myrand, dims, maxy, update)

from .gp import ModelOptimizer, NoModelOptimizer, MAPGPOptimizer, optimizemodel
from .acquisitionfunctions import (AbstractAcquisition, ProbabilityOfImprovement, ExpectedImprovement, UpperConfidenceBound,
                                   BrochuBetaScaling, NoBetaScaling, ThompsonSamplingSimple, MutualInformation, MaxMean,
                                   setparams, acquisitionfunction)
from .acquisition import defaultoptions, nlopt_setup, acquire_max, acquire_model_max
from .utils import (ScaledSobolIterator, ScaledLHSIterator, latin_hypercube_sampling, IterationCounter, DurationCounter)
from .bopt import (BOpt, boptimize, optimize, merge_with_defaults, maxduration, maxiterations, Sense, Verbosity, Min, Max,
                   Silent, Timings, Progress)
from . import dist

ElasticGPE = B200GPE    # drop-in name used by reference scripts (README.md:22-26)
