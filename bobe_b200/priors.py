"""Hyper-parameter priors on the host (O(d) work) with analytic gradients in log-parameter space.

The reference takes these from ``numpyro.distributions`` (not vendored; numpyro 0.15.x semantics restated):
``make_distribution`` BOBE/gp.py:27-54, ``saas_prior_logprob`` :56-78, ``_standard_prior_logprob`` :357-362,
prior set-up :309-337.  Note the reference's DEFAULT priors are Uniform over the bounds (constants with zero
gradient), not the LogNormal the docstring mentions (BOBE/gp.py:312-313,325-326).
"""
from __future__ import annotations

import math

import numpy as np

SQRT2 = math.sqrt(2.0)
SQRT3 = math.sqrt(3.0)


class Distribution:
    """log_prob(z) and d log_prob / d log z for a positive scalar/array z."""

    def log_prob(self, z):
        raise NotImplementedError

    def dlogp_dlogz(self, z):
        raise NotImplementedError


class Uniform(Distribution):
    def __init__(self, low=0.0, high=1.0):
        self.low, self.high = float(low), float(high)

    def log_prob(self, z):  # numpyro: constant, no support check unless validate_args
        return -math.log(self.high - self.low) * np.ones_like(np.asarray(z, dtype=np.float64))

    def dlogp_dlogz(self, z):
        return np.zeros_like(np.asarray(z, dtype=np.float64))


class LogNormal(Distribution):
    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = float(loc), float(scale)

    def log_prob(self, z):
        lz = np.log(np.asarray(z, dtype=np.float64))
        return -0.5 * ((lz - self.loc) / self.scale) ** 2 - math.log(self.scale * math.sqrt(2 * math.pi)) - lz

    def dlogp_dlogz(self, z):
        lz = np.log(np.asarray(z, dtype=np.float64))
        return -(lz - self.loc) / self.scale**2 - 1.0


class HalfCauchy(Distribution):
    def __init__(self, scale=1.0):
        self.scale = float(scale)

    def log_prob(self, z):
        z = np.asarray(z, dtype=np.float64)
        return math.log(2.0) - math.log(math.pi) - math.log(self.scale) - np.log1p((z / self.scale) ** 2)

    def dlogp_dlogz(self, z):
        u2 = (np.asarray(z, dtype=np.float64) / self.scale) ** 2
        return -2.0 * u2 / (1.0 + u2)


class Normal(Distribution):
    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = float(loc), float(scale)

    def log_prob(self, z):
        z = np.asarray(z, dtype=np.float64)
        return -0.5 * ((z - self.loc) / self.scale) ** 2 - math.log(self.scale * math.sqrt(2 * math.pi))

    def dlogp_dlogz(self, z):
        z = np.asarray(z, dtype=np.float64)
        return -(z - self.loc) / self.scale**2 * z


class Gamma(Distribution):
    def __init__(self, concentration=1.0, rate=1.0):
        self.a, self.b = float(concentration), float(rate)

    def log_prob(self, z):
        z = np.asarray(z, dtype=np.float64)
        return self.a * math.log(self.b) - math.lgamma(self.a) + (self.a - 1.0) * np.log(z) - self.b * z

    def dlogp_dlogz(self, z):
        z = np.asarray(z, dtype=np.float64)
        return (self.a - 1.0) - self.b * z


class DummyDistribution(Distribution):
    """BOBE/gp.py:22-25 -- used when the kernel variance is fixed."""

    def log_prob(self, z):
        return 0.0

    def dlogp_dlogz(self, z):
        return 0.0


_REGISTRY = {"Uniform": Uniform, "LogNormal": LogNormal, "HalfCauchy": HalfCauchy, "Normal": Normal, "Gamma": Gamma}


def make_distribution(spec: dict) -> Distribution:
    """BOBE/gp.py:27-54 -- dictionary spec {'name': ..., **params} -> distribution."""
    cls = _REGISTRY.get(spec["name"])
    if cls is None:
        raise ValueError(f"Distribution {spec['name']} not found in numpyro.distributions.")
    kwargs = {k: v for k, v in spec.items() if k != "name"}
    return cls(**kwargs)


def dslp_distribution(ndim: int) -> LogNormal:
    """BOBE/gp.py:330."""
    return LogNormal(loc=SQRT2 + 0.5 * math.log(ndim), scale=SQRT3)


def saas_prior_logprob(lengthscales, kernel_variance, tausq) -> float:
    """BOBE/gp.py:56-78 (no Jacobian terms; tausq only enters here, never the kernel)."""
    ls = np.asarray(lengthscales, dtype=np.float64)
    lp = float(LogNormal(0.0, 1.0).log_prob(kernel_variance))
    lp += float(HalfCauchy(0.1).log_prob(tausq))
    lp += float(np.sum(HalfCauchy(1.0).log_prob(1.0 / (tausq * ls**2))))
    return lp


def saas_prior_grad(lengthscales, kernel_variance, tausq):
    """d saas_prior_logprob / d(log l_j), d/d log kv, d/d log tausq."""
    ls = np.asarray(lengthscales, dtype=np.float64)
    u = 1.0 / (tausq * ls**2)
    w = u * u / (1.0 + u * u)
    g_ls = 4.0 * w  # u_j = 1/(tausq l_j^2): du/dlog l_j = -2u, dlog p/du = -2u/(1+u^2)
    g_kv = -math.log(kernel_variance) - 1.0
    t2 = (tausq / 0.1) ** 2
    g_tau = -2.0 * t2 / (1.0 + t2) + 2.0 * float(np.sum(w))
    return g_ls, g_kv, g_tau
