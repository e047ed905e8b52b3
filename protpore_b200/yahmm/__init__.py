"""Drop-in `yahmm` package (reference: yahmm/__init__.py:46 star-exports yahmm.pyx).

Same public names as the reference module: `Model`, `State`, the distributions, `log`, `exp`,
`log_probability`.  The dynamic-programming sweeps behind `Model` run on the GPU through the C ABI in
include/protpore_b200.h; see protpore_b200/yahmm/model.py.
"""
from .distributions import (Distribution, NormalDistribution, UniformDistribution, LogNormalDistribution,  # noqa: F401
                            ExponentialDistribution, LambdaDistribution, GaussianKernelDensity,
                            UniformKernelDensity, TriangleKernelDensity, MixtureDistribution, ExtremeValueDistribution,
                            GammaDistribution, InverseGammaDistribution, DiscreteDistribution, MultivariateDistribution)
from .state import State  # noqa: F401
from .model import Model, log, exp, log_probability  # noqa: F401

__version__ = '1.1.2'   # the reference's yahmm version (yahmm/__init__.py:48)
