"""Emission distributions of the drop-in `yahmm` API.

Mirrors the reference interface (yahmm.pyx:101-231 `Distribution`, 233-348 `UniformDistribution`,
350-523 `NormalDistribution`): same constructor arguments, `parameters` / `summaries` / `name` /
`frozen` attributes, and method names.  These Python methods are the *host-side* definition (used
for single-symbol queries, sampling and the Baum-Welch M-step); the per-observation evaluation on
the hot path happens inside the CUDA column kernels from the lowered (kind, p0, p1) table.

The two families the pro_001 peptide model uses are evaluated inside the kernels.  Every other family
(LogNormal, Exponential, Lambda, the Gaussian / Uniform / Triangle kernel densities PyPore/hmm.py and
PyPore/alignment.py use, Mixture, or any user subclass) is a TABLE state: the host fills the emission
matrix column e[:, k] = log_probability(x) for the batch, exactly what the reference does before every sweep
(yahmm.pyx:2836-2840), and the dynamic programming still runs on the GPU (exact state-parallel kernels).
"""
import math
import random

import numpy

NEGINF = float("-inf")
SQRT_2_PI = 2.50662827463   # the reference's truncated constant (yahmm.pyx:33) -- kept bit-for-bit

KIND_NORMAL = 0
KIND_UNIFORM = 1


def _log(x):
    """log(x) for x > 0, else -inf (yahmm.pyx:36-42)."""
    return math.log(x) if x > 0 else NEGINF


class Distribution(object):
    """Base class (yahmm.pyx:101-231)."""

    name = "Distribution"

    def __init__(self):
        self.name = "Distribution"
        self.frozen = False
        self.parameters = []
        self.summaries = []

    def __str__(self):
        parameters = [list(p) if isinstance(p, numpy.ndarray) else p for p in self.parameters]
        return "{}({})".format(self.name, ", ".join(map(str, parameters)))

    __repr__ = __str__

    def copy(self):
        return self.__class__(*self.parameters)

    def freeze(self):
        self.frozen = True

    def thaw(self):
        self.frozen = False

    def log_probability(self, symbol):
        raise NotImplementedError

    def sample(self):
        raise NotImplementedError

    def from_sample(self, items, weights=None):
        if self.frozen:
            return
        raise NotImplementedError

    def summarize(self, items, weights=None):
        # default: keep the raw sample (yahmm.pyx:198-218)
        if len(self.summaries) == 0:
            self.summaries = [items, weights]
        else:
            prior_items, prior_weights = self.summaries
            items = numpy.concatenate([prior_items, items])
            if weights is not None:
                weights = numpy.concatenate([prior_weights, weights])
            self.summaries = [items, weights]

    def from_summaries(self):
        if self.frozen:
            return
        self.from_sample(*self.summaries)
        self.summaries = []

    # ---- device lowering hook (new) -----------------------------------
    def device_params(self):
        """(kind, p0, p1) of the C ABI's emission description.  Normal and Uniform are evaluated inside the kernels;
        every other family is kind 2 (PP_EMIT_TABLE): the host evaluates `log_probability_many` over the batch's
        observations, as the reference fills its emission matrix (yahmm.pyx:2836-2840), and the kernels read the table."""
        return (2, 0.0, 0.0)

    def log_probability_many(self, symbols):
        """log_probability of every element of a 1-D array (families override this with a vectorised form)."""
        return numpy.array([self.log_probability(x) for x in symbols], dtype=numpy.float64)


class UniformDistribution(Distribution):
    """Uniform on [start, end] (yahmm.pyx:233-348)."""

    def __init__(self, start, end, frozen=False):
        self.parameters = [start, end]
        self.summaries = []
        self.name = "UniformDistribution"
        self.frozen = frozen

    def log_probability(self, symbol):
        a, b = float(self.parameters[0]), float(self.parameters[1])
        symbol = float(symbol)
        if symbol == a and symbol == b:
            return 0.0
        if a <= symbol <= b:
            return _log(1.0 / (b - a))
        return NEGINF

    def sample(self):
        return random.uniform(self.parameters[0], self.parameters[1])

    def from_sample(self, items, weights=None, inertia=0.0):
        if self.frozen:
            return
        weights = numpy.ones_like(items) if weights is None else numpy.asarray(weights)
        if weights.sum() == 0 or len(items) == 0:
            return
        lo, hi = self.parameters
        self.parameters[0] = lo * inertia + numpy.min(items) * (1 - inertia)
        self.parameters[1] = hi * inertia + numpy.max(items) * (1 - inertia)

    def summarize(self, items, weights=None):
        # the summary is [min, max] of the WHOLE sample whenever the total weight is non-zero (304-326)
        weights = numpy.ones_like(items) if weights is None else numpy.asarray(weights)
        if weights.sum() == 0 or len(items) == 0:
            return
        items = numpy.asarray(items)
        self.summaries.append([items.min(), items.max()])

    def summarize_range(self, w_sum, lo, hi):
        """Batched-path entry (new): one summary from the allreduced statistics of the GPU E-step -- the min / max over
        every sequence that gave the state non-zero weight.  The range itself says whether there was such a sequence
        (the kernels decide that per event, as the reference's `weights.sum() != 0` does, 304-326); the total weight is
        no substitute: a sum of weights around 1e-300 is not representable next to the scaled sweep's other terms."""
        if not (lo <= hi):
            return
        self.summaries.append([lo, hi])

    def from_summaries(self, inertia=0.0):
        if self.frozen:
            return
        if len(self.summaries) == 0:
            # the reference indexes s[:, 0] unguarded here (338-348) and raises IndexError when a
            # Uniform state got no weight from any sequence; we keep the parameters instead.
            return
        s = numpy.asarray(self.summaries)
        lo, hi = self.parameters
        self.parameters = [lo * inertia + s[:, 0].min() * (1 - inertia),
                           hi * inertia + s[:, 1].max() * (1 - inertia)]
        self.summaries = []

    def device_params(self):
        return KIND_UNIFORM, float(self.parameters[0]), float(self.parameters[1])


class NormalDistribution(Distribution):
    """Normal(mean, std) with the sigma == 0 point-mass branch (yahmm.pyx:350-523)."""

    def __init__(self, mean, std, frozen=False):
        self.parameters = [mean, std]
        self.summaries = []
        self.name = "NormalDistribution"
        self.frozen = frozen

    def log_probability(self, symbol, epsilon=1E-4):
        mu, sigma = float(self.parameters[0]), float(self.parameters[1])
        symbol = float(symbol)
        if sigma == 0.0:
            return 0.0 if abs(symbol - mu) < epsilon else NEGINF
        d = symbol - mu
        # C evaluates pow(d, 2) as d*d and 2*sigma**2 as 2*(sigma*sigma) (yahmm.pyx:389-390)
        return _log(1.0 / (sigma * SQRT_2_PI)) - (d * d) / (2 * (sigma * sigma))

    def sample(self):
        return random.normalvariate(*self.parameters)

    def from_sample(self, items, weights=None, inertia=0.0, min_std=0.01):
        if len(items) == 0 or self.frozen:
            return
        items = numpy.asarray(items)
        weights = numpy.ones_like(items) if weights is None else numpy.asarray(weights)
        if weights.sum() == 0:
            return
        mean = numpy.average(items, weights=weights)
        if len(weights[weights != 0]) > 1:
            variance = numpy.dot(items ** 2 - mean ** 2, weights) / weights.sum()
            std = math.sqrt(variance) if variance >= 0 else 0
        else:
            std = self.parameters[1]
        std = max(numpy.array([std, min_std]))
        prior_mean, prior_std = self.parameters
        self.parameters = [prior_mean * inertia + mean * (1 - inertia),
                           prior_std * inertia + std * (1 - inertia)]

    def summarize(self, items, weights=None):
        items = numpy.asarray(items)
        weights = numpy.ones_like(items) if weights is None else numpy.asarray(weights)
        if weights.sum() == 0:
            return
        mean = numpy.average(items, weights=weights)
        variance = numpy.dot(items ** 2 - mean ** 2, weights) / weights.sum()
        self.summaries.append([mean, variance, weights.sum()])

    def summarize_moments(self, w_sum, wx_sum, wxx_sum):
        """Batched-path entry (new): push one summary from global moments (sum w, sum w x, sum w x^2).

        Algebraically the same triple `summarize` stores ([mean, var, sum w], 462-486); lets the GPU
        E-step hand over allreduced sufficient statistics instead of per-sequence weight vectors.
        """
        if w_sum == 0:
            return
        mean = wx_sum / w_sum
        self.summaries.append([mean, wxx_sum / w_sum - mean * mean, w_sum])

    def from_summaries(self, inertia=0.0, min_std=0.01):
        if len(self.summaries) == 0 or self.frozen:
            return
        s = numpy.asarray(self.summaries)
        mean = numpy.average(s[:, 0], weights=s[:, 2])
        variance = numpy.sum([(v + m ** 2) * w for m, v, w in s]) / s[:, 2].sum() - mean ** 2
        std = math.sqrt(variance) if variance >= 0 else 0
        std = max(min_std, std)
        prior_mean, prior_std = self.parameters
        self.parameters = [prior_mean * inertia + mean * (1 - inertia),
                           prior_std * inertia + std * (1 - inertia)]
        self.summaries = []

    def device_params(self):
        return KIND_NORMAL, float(self.parameters[0]), float(self.parameters[1])


# ---------------------------------------------------------------------------------------------------------------
# Families the kernels read from the emission table.  Same constructor arguments, `parameters`,
# `log_probability`, `sample` and `from_sample` as the reference; no device lowering of their own.
# ---------------------------------------------------------------------------------------------------------------
class LogNormalDistribution(Distribution):
    """yahmm.pyx:525-688."""

    def __init__(self, mu, sigma, frozen=False):
        Distribution.__init__(self)
        self.parameters = [mu, sigma]
        self.name = "LogNormalDistribution"
        self.frozen = frozen

    def log_probability(self, symbol):
        mu, sigma = self.parameters                                            # 547-556
        return -math.log(symbol * sigma * SQRT_2_PI) - 0.5 * ((math.log(symbol) - mu) / sigma) ** 2

    def sample(self):
        return numpy.random.lognormal(*self.parameters)

    def from_sample(self, items, weights=None, inertia=0.0, min_std=0.01):
        if len(items) == 0 or self.frozen:
            return
        items = numpy.asarray(items)
        weights = numpy.ones_like(items) if weights is None else numpy.asarray(weights)
        if weights.sum() == 0:
            return
        logs = numpy.log(items)
        mean = numpy.average(logs, weights=weights)                            # 595-613
        if len(weights[weights != 0]) > 1:
            variance = numpy.dot(logs ** 2 - mean ** 2, weights) / weights.sum()
            std = math.sqrt(variance) if variance >= 0 else 0                  # a small negative variance by accident: 0
        else:
            std = self.parameters[1]                                           # one data point cannot update the std
        std = max(std, min_std)
        prior_mean, prior_std = self.parameters
        self.parameters = [prior_mean * inertia + mean * (1 - inertia), prior_std * inertia + std * (1 - inertia)]

    def summarize(self, items, weights=None):
        # (mean of the logs, variance, weight) per call (626-648)
        items = numpy.asarray(items)
        weights = numpy.ones_like(items) if weights is None else numpy.asarray(weights)
        if weights.sum() == 0:
            return
        logs = numpy.log(items)
        mean = numpy.average(logs, weights=weights)
        variance = numpy.dot(logs ** 2 - mean ** 2, weights) / weights.sum()
        self.summaries.append([mean, variance, weights.sum()])

    def from_summaries(self, inertia=0.0, min_std=0.01):
        # merged mean and variance of the summaries (653-688)
        if len(self.summaries) == 0 or self.frozen:
            return
        summaries = numpy.asarray(self.summaries)
        mean = numpy.average(summaries[:, 0], weights=summaries[:, 2])
        variance = numpy.sum([(v + m ** 2) * w for m, v, w in summaries]) / summaries[:, 2].sum() - mean ** 2
        std = max(min_std, math.sqrt(variance) if variance >= 0 else 0)
        prior_mean, prior_std = self.parameters
        self.parameters = [prior_mean * inertia + mean * (1 - inertia), prior_std * inertia + std * (1 - inertia)]
        self.summaries = []


class ExponentialDistribution(Distribution):
    """yahmm.pyx:726-842."""

    def __init__(self, rate, frozen=False):
        Distribution.__init__(self)
        self.parameters = [rate]
        self.name = "ExponentialDistribution"
        self.frozen = frozen

    def log_probability(self, symbol):
        return _log(self.parameters[0]) - self.parameters[0] * symbol          # 741-746

    def log_probability_many(self, symbols):
        return _log(self.parameters[0]) - self.parameters[0] * numpy.asarray(symbols, dtype=numpy.float64)

    def sample(self):
        return random.expovariate(*self.parameters)

    def from_sample(self, items, weights=None, inertia=0.0):
        if len(items) == 0 or self.frozen:
            return
        items = numpy.asarray(items)
        weights = numpy.ones_like(items) if weights is None else numpy.asarray(weights)
        if weights.sum() == 0:
            return
        rate = 1.0 / numpy.average(items, weights=weights)
        self.parameters[0] = self.parameters[0] * inertia + rate * (1 - inertia)

    def summarize(self, items, weights=None):
        # (mean, weight) per call (796-815)
        items = numpy.asarray(items)
        weights = numpy.ones_like(items) if weights is None else numpy.asarray(weights)
        if weights.sum() == 0:
            return
        self.summaries.append([numpy.average(items, weights=weights), weights.sum()])

    def from_summaries(self, inertia=0.0):
        # 817-842
        if len(self.summaries) == 0 or self.frozen:
            return
        summaries = numpy.asarray(self.summaries)
        rate = 1.0 / numpy.average(summaries[:, 0], weights=summaries[:, 1])
        self.parameters[0] = self.parameters[0] * inertia + rate * (1 - inertia)
        self.summaries = []


class LambdaDistribution(Distribution):
    """A log-density given as a Python callable (yahmm.pyx:1326-1356)."""

    def __init__(self, lambda_funct, frozen=True):
        Distribution.__init__(self)
        self.parameters = [lambda_funct]
        self.name = "LambdaDistribution"
        self.frozen = frozen

    def log_probability(self, symbol):
        return self.parameters[0](symbol)


class _KernelDensity(Distribution):
    """Points, bandwidth, weights (normalised to sum 1); shared by the three kernel densities."""

    def __init__(self, points, bandwidth=1, weights=None, frozen=False):
        Distribution.__init__(self)
        points = numpy.asarray(points)
        n = len(points)
        if weights is not None and len(weights) > 0:
            weights = numpy.array(weights) / numpy.sum(weights)
        else:
            weights = numpy.ones(n) / n
        self.parameters = [points, bandwidth, weights]
        self.frozen = frozen

    def _kernel(self, diff):
        raise NotImplementedError

    def log_probability(self, symbol):
        # sequential accumulation over the points in stored order, as the reference's loop (1395-1416)
        total = 0.0
        points, _, weights = self.parameters
        contrib = self._kernel(points - symbol)
        for c, w in zip(contrib, weights):
            total += c * w
        return _log(total)

    def log_probability_many(self, symbols):
        return numpy.array([self.log_probability(float(x)) for x in symbols], dtype=numpy.float64)

    def from_sample(self, points, weights=None, inertia=0.0):
        if self.frozen:
            return
        points = numpy.asarray(points)
        n = len(points)
        if weights is not None and len(weights) > 0:
            weights = numpy.array(weights) / numpy.sum(weights)
        else:
            weights = numpy.ones(n) / n
        if inertia == 0.0:
            self.parameters[0] = points
            self.parameters[2] = weights
        else:
            self.parameters[0] = numpy.concatenate((self.parameters[0], points))
            self.parameters[2] = numpy.concatenate((self.parameters[2] * inertia, weights * (1 - inertia)))


class GaussianKernelDensity(_KernelDensity):
    """yahmm.pyx:1358-1456: log sum_i w_i exp(-0.5 ((p_i - x) / bandwidth)^2) / sqrt(2 pi)."""

    def __init__(self, points, bandwidth=1, weights=None, frozen=False):
        _KernelDensity.__init__(self, points, bandwidth, weights, frozen)
        self.name = "GaussianKernelDensity"

    def _kernel(self, diff):
        return (1.0 / SQRT_2_PI) * numpy.exp(-0.5 * (diff / self.parameters[1]) ** 2)

    def sample(self):
        mu = numpy.random.choice(self.parameters[0], p=self.parameters[2])
        return random.gauss(mu, self.parameters[1])


class UniformKernelDensity(_KernelDensity):
    """yahmm.pyx:1458-1559: a point counts 1 when |p_i - x| <= bandwidth."""

    def __init__(self, points, bandwidth=1, weights=None, frozen=False):
        _KernelDensity.__init__(self, points, bandwidth, weights, frozen)
        self.name = "UniformKernelDensity"

    def _kernel(self, diff):
        return (numpy.abs(diff) <= self.parameters[1]).astype(numpy.float64)

    def sample(self):
        mu = numpy.random.choice(self.parameters[0], p=self.parameters[2])
        return random.uniform(mu - self.parameters[1], mu + self.parameters[1])


class TriangleKernelDensity(_KernelDensity):
    """yahmm.pyx:1561-1660: max(0, bandwidth - |p_i - x|) per point."""

    def __init__(self, points, bandwidth=1, weights=None, frozen=False):
        _KernelDensity.__init__(self, points, bandwidth, weights, frozen)
        self.name = "TriangleKernelDensity"

    def _kernel(self, diff):
        return numpy.maximum(self.parameters[1] - numpy.abs(diff), 0.0)

    def sample(self):
        mu = numpy.random.choice(self.parameters[0], p=self.parameters[2])
        return random.triangular(mu - self.parameters[1], mu + self.parameters[1], mu)


class MixtureDistribution(Distribution):
    """Weighted mixture of other distributions (yahmm.pyx:1662-1821)."""

    def __init__(self, distributions, weights=None, frozen=False):
        Distribution.__init__(self)
        n = len(distributions)
        if weights is not None and len(weights) > 0:
            weights = numpy.array(weights) / numpy.sum(weights)
        else:
            weights = numpy.ones(n) / n
        self.parameters = [distributions, weights]
        self.name = "MixtureDistribution"
        self.frozen = frozen

    def log_probability(self, symbol):
        d, w = self.parameters                                                 # 1704-1706
        return _log(numpy.sum([math.exp(d[i].log_probability(symbol)) * w[i] for i in range(len(d))]))

    def sample(self):
        i = random.random()
        for d, w in zip(*self.parameters):
            if w > i:
                return d.sample()
            i -= w


# ---- the remaining families of the reference ---------------------------------------------------
# Not used by any model builder in the repository.  They evaluate their density exactly as the reference does and reach
# the GPU sweeps through the emission-table path like every family the kernels do not evaluate themselves
# (device_params() -> PP_EMIT_TABLE).  Parameter re-estimation is implemented where the reference's is a closed form
# (Discrete); Gamma / InverseGamma / ExtremeValue keep their parameters under training unless thawed by a user subclass
# (the reference's Gamma fit is an iterative scheme built on scipy.special, yahmm.pyx:886-1119).

class ExtremeValueDistribution(Distribution):
    """Generalised extreme value distribution (yahmm.pyx:690-724); frozen by default as in the reference."""

    def __init__(self, mu, sigma, epsilon, frozen=True):
        self.parameters = [float(mu), float(sigma), float(epsilon)]
        self.summaries = []
        self.name = "ExtremeValueDistribution"
        self.frozen = frozen

    def log_probability(self, symbol):
        mu, sigma, epsilon = self.parameters
        t = (float(symbol) - mu) / sigma
        if epsilon == 0:
            return -math.log(sigma) - t - math.exp(-t)
        return -math.log(sigma) + math.log(1 + epsilon * t) * (-1. / epsilon - 1) - (1 + epsilon * t) ** (-1. / epsilon)


class GammaDistribution(Distribution):
    """Gamma distribution, shape / rate parameterisation (yahmm.pyx:844-1119)."""

    def __init__(self, alpha, beta, frozen=False):
        self.parameters = [alpha, beta]
        self.summaries = []
        self.name = "GammaDistribution"
        self.frozen = frozen

    def log_probability(self, symbol):
        alpha, beta = self.parameters
        return _log(beta) * alpha - math.lgamma(alpha) + _log(symbol) * (alpha - 1) - beta * symbol

    def sample(self):
        return random.gammavariate(self.parameters[0], 1.0 / self.parameters[1])

    def from_sample(self, items, weights=None, inertia=0.0):
        # the reference fits by an iterative scheme on scipy.special (886-1023); here the method of moments, which is
        # that scheme's starting point (943-951)
        if self.frozen or len(items) == 0:
            return
        items = numpy.asarray(items, dtype=numpy.float64)
        weights = numpy.ones_like(items) if weights is None else numpy.asarray(weights, dtype=numpy.float64)
        if weights.sum() == 0:
            return
        mean = numpy.average(items, weights=weights)
        var = numpy.average((items - mean) ** 2, weights=weights)
        if var <= 0:
            return
        alpha, beta = mean * mean / var, mean / var
        self.parameters = [self.parameters[0] * inertia + alpha * (1 - inertia), self.parameters[1] * inertia + beta * (1 - inertia)]


class InverseGammaDistribution(GammaDistribution):
    """1 / X ~ Gamma(alpha, beta) (yahmm.pyx:1121-1184)."""

    def __init__(self, alpha, beta, frozen=False):
        GammaDistribution.__init__(self, alpha, beta, frozen)
        self.name = "InverseGammaDistribution"

    def log_probability(self, symbol):
        return GammaDistribution.log_probability(self, 1.0 / symbol)

    def sample(self):
        return 1.0 / GammaDistribution.sample(self)

    def from_sample(self, items, weights=None, inertia=0.0):
        GammaDistribution.from_sample(self, 1.0 / numpy.asarray(items, dtype=numpy.float64), weights, inertia)


class DiscreteDistribution(Distribution):
    """Characters (here: the float values an observation can take) and their probabilities (yahmm.pyx:1186-1330)."""

    def __init__(self, characters, frozen=False):
        self.parameters = [characters]
        self.summaries = [{}, 0]
        self.name = "DiscreteDistribution"
        self.frozen = frozen

    def log_probability(self, symbol):
        return _log(self.parameters[0].get(symbol, 0))

    def sample(self):
        rand = random.random()
        for key, value in self.parameters[0].items():
            if value >= rand:
                return key
            rand -= value

    def from_sample(self, items, weights=None, inertia=0.0):
        if self.frozen or len(items) == 0:
            return
        weights = numpy.ones(len(items)) if weights is None else numpy.asarray(weights, dtype=numpy.float64)
        if weights.sum() == 0:
            return
        counts = {}
        for x, w in zip(items, weights):
            counts[x] = counts.get(x, 0.0) + w
        total = sum(counts.values())
        prior = self.parameters[0]
        keys = set(prior) | set(counts)
        self.parameters = [{k: prior.get(k, 0.0) * inertia + counts.get(k, 0.0) / total * (1 - inertia) for k in keys}]

    def summarize(self, items, weights=None):
        weights = numpy.ones(len(items)) if weights is None else numpy.asarray(weights, dtype=numpy.float64)
        for x, w in zip(items, weights):
            self.summaries[0][x] = self.summaries[0].get(x, 0.0) + w
        self.summaries[1] += float(weights.sum())

    def from_summaries(self, inertia=0.0):
        if self.frozen or self.summaries[1] == 0:
            return
        counts, total = self.summaries
        prior = self.parameters[0]
        keys = set(prior) | set(counts)
        self.parameters = [{k: prior.get(k, 0.0) * inertia + counts.get(k, 0.0) / total * (1 - inertia) for k in keys}]
        self.summaries = [{}, 0]


class MultivariateDistribution(Distribution):
    """Independent components over tuple-valued observations (yahmm.pyx:1823-1917).  The batched GPU calls take scalar
    observations; a model with such a state is served symbol by symbol through `log_probability` on the host."""

    def __init__(self, distributions, weights=None, frozen=False):
        n = len(distributions)
        weights = numpy.array(weights) if weights else numpy.ones(n)
        self.parameters = [distributions, weights]
        self.summaries = []
        self.name = "MultivariateDistribution"
        self.frozen = frozen

    def __str__(self):
        return "MultivariateDistribution({})".format([str(d) for d in self.parameters[0]]).replace("'", "")

    def log_probability(self, symbol):
        return sum(d.log_probability(obs) * w for d, obs, w in zip(self.parameters[0], symbol, self.parameters[1]))

    def sample(self):
        return [d.sample() for d in self.parameters[0]]

    def from_sample(self, items, weights=None, inertia=0.0):
        if self.frozen:
            return
        items = numpy.asarray(items)
        for i, d in enumerate(self.parameters[0]):
            d.from_sample(items[:, i], weights=weights, inertia=inertia)
