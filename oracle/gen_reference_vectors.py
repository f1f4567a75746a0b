"""Golden vectors produced by the REFERENCE'S OWN SOURCE (test infrastructure; runs only where /root/reference exists).

JAX, numpyro, optax and tensorflow-probability are not installed in this image and cannot be, so ``import BOBE`` fails.
This script therefore executes the reference's unmodified source files ``BOBE/gp.py``, ``BOBE/acquisition.py`` (and the
helpers they import: ``optim.py``, ``utils/core.py``, ``utils/log.py``, ``utils/seed.py``) from where they lie under
/root/reference, with a NumPy/SciPy stand-in for the part of the jax API those files touch:

    jax.numpy            -> numpy (array creators return a subclass that adds ``.at[idx].set(v)``)
    jax.scipy.linalg     -> scipy.linalg (cho_solve, solve_triangular: the same LAPACK routines XLA's CPU backend calls)
    jax.scipy.stats.norm -> scipy.stats.norm
    jax.jit              -> identity;   jax.vmap / lax.map -> a Python loop over the leading axis
    jax.value_and_grad   -> not available under NumPy; BOBE/gp.py is therefore loaded a SECOND time with jax.numpy ->
                            torch float64 tensors (load_reference_autodiff), where jax.value_and_grad(gp.neg_mll) runs as
                            reverse-mode autodiff through the reference's own statements
    numpyro.distributions.Uniform.log_prob -> -log(high - low)   (the only prior the default GP constructs)
    tensorflow_probability...math.erfcx / log1mexp -> scipy.special.erfcx / log(1 - exp(-|x|)) (TFP's documented definition)

What this pins: every ARITHMETIC STATEMENT of the reference on the path (distances, both kernels, the log marginal
likelihood, the rank-1 Cholesky append, standardisation, Cholesky + alphas, posterior mean / variance in all their
variants, duplicate handling in ``update``, the fantasy variance, EI / LogEI / WIPV / WIPStd values) is executed as written
by the reference's authors, in float64, on seeded inputs; the outputs are stored next to the inputs in
``tests/golden/reference_source_vectors.npz``.  ``tests/test_oracle.py`` checks the oracle restatement against them and
``tests/test_gpu_parity.py`` checks the CUDA path against them.
The log-ML gradient -- what ``jax.value_and_grad`` gives the optimisers at BOBE/optim.py:118,211,309 -- comes from the
torch-backed load (``neg_mll_ad_grad``), with central differences of the NumPy-run ``neg_mll`` stored beside it; the same
load runs the reference's ``GP.fit`` -> ``optimize_scipy`` and differentiates ``predict_single``, ``EI.fun``, ``LogEI.fun``,
``WIPV.fun`` and ``WIPStd.fun`` in the query / candidate point.
What it does NOT pin: XLA's own floating-point behaviour (fusion, its Cholesky kernel and that kernel's derivative rule) --
rounding-level differences -- and the numpyro priors other than Uniform.

    python oracle/gen_reference_vectors.py            # writes tests/golden/reference_source_vectors.npz
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import scipy.linalg
import scipy.special
import scipy.stats

REF = os.environ.get("BOBE_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "reference_source_vectors.npz")


# ---- the stand-in for the jax API -----------------------------------------------------------------------------------
class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Ref:
            def set(self, v):
                out = np.array(arr, copy=True).view(JArray)
                out[idx] = v
                return out

            def add(self, v):
                out = np.array(arr, copy=True).view(JArray)
                out[idx] += v
                return out
        return _Ref()


class JArray(np.ndarray):
    """numpy array with jax's functional-update syntax ``x.at[idx].set(v)``."""
    @property
    def at(self):
        return _At(self)


def _wrap(fn):
    def inner(*a, **k):
        r = fn(*a, **k)
        return r.view(JArray) if isinstance(r, np.ndarray) else r
    inner.__name__ = getattr(fn, "__name__", "wrapped")
    return inner


def _make_jnp():
    m = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("_"):
            setattr(m, name, getattr(np, name))
    for name in ("zeros", "ones", "array", "asarray", "eye", "empty", "full", "zeros_like", "ones_like", "arange", "linspace"):
        setattr(m, name, _wrap(getattr(np, name)))

    def clip(a, a_min=None, a_max=None, **kw):  # jnp.clip(x, a_min=...) with either bound optional
        a_min = kw.pop("min", a_min)
        a_max = kw.pop("max", a_max)
        return np.clip(a, a_min, a_max)
    m.clip = clip
    m.ndarray = np.ndarray
    m.linalg = np.linalg
    return m


def _jit(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


def _vmap(fn, in_axes=0, out_axes=0):
    if in_axes != 0:
        raise NotImplementedError("stand-in vmap: only in_axes=0 is used on this path")

    def mapped(xs):
        outs = [fn(x) for x in xs]
        if isinstance(outs[0], tuple):
            return tuple(np.stack([np.asarray(o[i]) for o in outs]) for i in range(len(outs[0])))
        return np.stack([np.asarray(o) for o in outs])
    return mapped


def _no_autodiff(*_a, **_k):
    raise NotImplementedError("jax autodiff is not available in the NumPy stand-in")


def install_standin():
    jnp = _make_jnp()
    jax = types.ModuleType("jax")
    jax.__path__ = []
    jax.numpy = jnp
    jax.jit = _jit
    jax.vmap = _vmap
    jax.value_and_grad = _no_autodiff
    jax.grad = _no_autodiff
    jax.Array = np.ndarray
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    lax = types.ModuleType("jax.lax")
    lax.map = lambda f, xs, **_k: _vmap(f)(xs)
    jax.lax = lax
    rnd = types.ModuleType("jax.random")
    rnd.PRNGKey = lambda seed: np.array([0, seed], dtype=np.uint32)
    rnd.split = lambda key, num=2: [np.array([int(key[1]) + i + 1, int(key[1])], dtype=np.uint32) for i in range(num)]
    jax.random = rnd
    jsp = types.ModuleType("jax.scipy")
    jsp.__path__ = []
    jsl = types.ModuleType("jax.scipy.linalg")
    jsl.cho_solve = scipy.linalg.cho_solve
    jsl.solve_triangular = scipy.linalg.solve_triangular
    jss = types.ModuleType("jax.scipy.stats")
    jss.norm = scipy.stats.norm
    jsx = types.ModuleType("jax.scipy.special")
    for name in ("erfcx", "erfc", "erf", "ndtr", "log_ndtr", "logsumexp", "gammaln"):
        setattr(jsx, name, getattr(scipy.special, name))
    jsp.linalg, jsp.stats, jsp.special = jsl, jss, jsx
    jax.scipy = jsp
    mods = {"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.random": rnd, "jax.scipy": jsp, "jax.scipy.linalg": jsl,
            "jax.scipy.stats": jss, "jax.scipy.special": jsx}

    # numpyro.distributions: only what the default GP constructs (Uniform priors: a constant)
    numpyro = types.ModuleType("numpyro")
    numpyro.__path__ = []
    dist = types.ModuleType("numpyro.distributions")

    class Distribution:
        pass

    class Uniform(Distribution):
        def __init__(self, low=0.0, high=1.0):
            self.low, self.high = low, high

        def log_prob(self, x):
            return -np.log(self.high - self.low) + 0.0 * np.asarray(x)

    # the other priors the reference names (DSLP, SAAS, user specs): the densities come from scipy.stats -- third-party
    # code for the same named distributions -- so what these cases pin is the reference's COMPOSITION of them (which
    # argument meets which density, 1 / (tausq l^2), the DSLP location sqrt2 + log(d) / 2, ...)
    class LogNormal(Distribution):
        def __init__(self, loc=0.0, scale=1.0):
            self.loc, self.scale = float(loc), float(scale)

        def log_prob(self, x):
            return scipy.stats.lognorm.logpdf(np.asarray(x, dtype=np.float64), s=self.scale, scale=np.exp(self.loc))

    class HalfCauchy(Distribution):
        def __init__(self, scale=1.0):
            self.scale = float(scale)

        def log_prob(self, x):
            return scipy.stats.halfcauchy.logpdf(np.asarray(x, dtype=np.float64), scale=self.scale)

    class Normal(Distribution):
        def __init__(self, loc=0.0, scale=1.0):
            self.loc, self.scale = float(loc), float(scale)

        def log_prob(self, x):
            return scipy.stats.norm.logpdf(np.asarray(x, dtype=np.float64), loc=self.loc, scale=self.scale)
    dist.Distribution, dist.Uniform, dist.LogNormal, dist.HalfCauchy, dist.Normal = Distribution, Uniform, LogNormal, HalfCauchy, Normal
    numpyro.distributions = dist
    mods.update({"numpyro": numpyro, "numpyro.distributions": dist})

    # tensorflow_probability.substrates.jax: tfp.math.erfcx, tfp.math.log1mexp
    tfp_root = types.ModuleType("tensorflow_probability")
    tfp_root.__path__ = []
    sub = types.ModuleType("tensorflow_probability.substrates")
    sub.__path__ = []
    tfj = types.ModuleType("tensorflow_probability.substrates.jax")

    def log1mexp(x):  # TFP: log(1 - exp(-|x|)), Maechler's two-branch evaluation
        x = np.abs(np.asarray(x, dtype=np.float64))
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(x < np.log(2.0), np.log(-np.expm1(-x)), np.log1p(-np.exp(-x)))
    tfj.math = types.SimpleNamespace(erfcx=scipy.special.erfcx, log1mexp=log1mexp)
    tfp_root.substrates, sub.jax = sub, tfj
    mods.update({"tensorflow_probability": tfp_root, "tensorflow_probability.substrates": sub,
                 "tensorflow_probability.substrates.jax": tfj})
    sys.modules.update(mods)


def load_reference():
    """Import BOBE.gp and BOBE.acquisition from their source files without running BOBE/__init__.py (which pulls in the
    samplers, plotting and MPI layers)."""
    install_standin()
    pkg = types.ModuleType("BOBE")
    pkg.__path__ = [os.path.join(REF, "BOBE")]
    utils = types.ModuleType("BOBE.utils")
    utils.__path__ = [os.path.join(REF, "BOBE", "utils")]
    samplers = types.ModuleType("BOBE.samplers")  # acquisition.py imports two names from it at module level
    samplers.nested_sampling_Dy = samplers.sample_GP_NUTS = None
    sys.modules.update({"BOBE": pkg, "BOBE.utils": utils, "BOBE.samplers": samplers})

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "BOBE", rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    load("BOBE.utils.log", "utils/log.py")
    load("BOBE.utils.seed", "utils/seed.py")
    load("BOBE.utils.core", "utils/core.py")
    load("BOBE.optim", "optim.py")
    gp = load("BOBE.gp", "gp.py")
    acq = load("BOBE.acquisition", "acquisition.py")
    clf = load("BOBE.clf", "clf.py")  # (scikit-learn's SVC is installed; flax / optax are absent: SVM only, as in the product)
    clf_gp = load("BOBE.clf_gp", "clf_gp.py")
    clf_gp.svm_predict = clf.svm_predict
    return gp, acq, clf_gp


# ---- a second stand-in, backed by torch, for the one thing NumPy cannot do: jax.value_and_grad ------------------------
def load_reference_autodiff():
    """BOBE/gp.py loaded a second time (package name BOBE_ad) with jax.numpy -> torch float64 tensors, so that
    ``jax.value_and_grad(gp.neg_mll)`` -- what BOBE/optim.py:118,211,309 differentiates -- runs as reverse-mode autodiff
    through the reference's own statements (kernel, Cholesky, cho_solve, log-determinant)."""
    import math

    import torch

    f64 = torch.float64

    def as_t(x, dtype=None):
        if isinstance(x, torch.Tensor):
            return x
        return torch.as_tensor(np.asarray(x, dtype=np.float64) if not isinstance(x, (int, float)) else x, dtype=f64)

    jnp = types.ModuleType("jax.numpy")
    jnp.ndarray, jnp.float64, jnp.float32, jnp.pi = torch.Tensor, f64, torch.float32, math.pi
    jnp.array = jnp.asarray = as_t
    class TAt:
        def __init__(self, t):
            self.t = t

        def __getitem__(self, idx):
            t = self.t

            class _Ref:
                def set(self, v):
                    out = t.clone()
                    out[idx] = v
                    return out.as_subclass(TArray)
            return _Ref()

    class TArray(torch.Tensor):
        """torch tensor with jax's functional update ``x.at[idx].set(v)`` (differentiable: clone + index_put)."""
        @property
        def at(self):
            return TAt(self)

    jnp.ones = lambda *shape, **k: torch.ones(*shape, dtype=f64)
    jnp.zeros = lambda *shape, **k: torch.zeros(*shape, dtype=f64).as_subclass(TArray)
    jnp.full = lambda shape, v, **k: as_t(v).expand(shape).clone() if isinstance(v, torch.Tensor) else torch.full(shape, v, dtype=f64)
    jnp.reshape = lambda x, shape: torch.reshape(x, shape)
    jnp.argmax = lambda x, axis=None: torch.argmax(as_t(x))
    jnp.argmin = lambda x, axis=None: torch.argmin(as_t(x))
    jnp.min = lambda x, axis=None: torch.min(as_t(x))
    jnp.max = lambda x, axis=None: torch.max(as_t(x))
    jnp.eye = lambda n, **k: torch.eye(n, dtype=f64)
    for name in ("exp", "log", "sqrt", "square", "abs", "isnan", "diag", "dot", "vstack", "hstack", "concatenate", "isclose"):
        setattr(jnp, name, (lambda fn: (lambda *a, **k: fn(*[as_t(x) if not isinstance(x, (list, tuple)) else [as_t(e) for e in x]
                                                             for x in a], **k)))(
            getattr(torch, name if name != "concatenate" else "cat")))
    jnp.tile = lambda x, reps: torch.tile(as_t(x), reps)
    jnp.sum = lambda x, axis=None: torch.sum(as_t(x)) if axis is None else torch.sum(as_t(x), dim=axis)
    # (standardisation constants of NumPy training data stay NumPy scalars: ``ndarray - Tensor`` is not defined)
    jnp.mean = lambda x, axis=None: float(np.mean(x)) if isinstance(x, np.ndarray) else torch.mean(x)
    jnp.std = lambda x, axis=None: float(np.std(x)) if isinstance(x, np.ndarray) else torch.std(x, correction=0)
    jnp.any, jnp.all = (lambda x, axis=None: torch.any(x) if axis is None else torch.any(x, dim=axis)), \
        (lambda x, axis=None: torch.all(x) if axis is None else torch.all(x, dim=axis))
    jnp.where = lambda c, a, b: torch.where(c, as_t(a), as_t(b))
    jnp.clip = lambda a, a_min=None, a_max=None: torch.clamp(as_t(a), min=a_min, max=a_max)
    jnp.atleast_2d = lambda x: torch.atleast_2d(as_t(x))
    jnp.einsum = lambda spec, *ops: torch.einsum(spec if "->" in spec else spec + "->", *[as_t(o) for o in ops])
    jnp.linalg = types.SimpleNamespace(cholesky=torch.linalg.cholesky)

    def cho_solve(c_and_lower, b):
        L, lower = c_and_lower
        b = as_t(b)
        return torch.cholesky_solve(b if b.ndim == 2 else b[:, None], L, upper=not lower)

    def solve_triangular(a, b, lower=False):
        b2 = b if b.ndim == 2 else b[:, None]
        out = torch.linalg.solve_triangular(a, b2, upper=not lower)
        return out if b.ndim == 2 else out[:, 0]

    def value_and_grad(fn):
        def vg(x, *a, **k):
            xt = torch.tensor(np.asarray(x, dtype=np.float64), dtype=f64, requires_grad=True)
            val = fn(xt, *a, **k)
            (g,) = torch.autograd.grad(val, xt)
            return float(val.detach()), g.numpy().copy()
        return vg

    def tvmap(fn, in_axes=0, out_axes=0):
        def mapped(xs):
            outs = [fn(x) for x in xs]
            if isinstance(outs[0], tuple):
                return tuple(torch.stack([o[i] for o in outs]) for i in range(len(outs[0])))
            return torch.stack(outs)
        return mapped

    jax = types.ModuleType("jax")
    jax.__path__ = []
    jax.numpy, jax.jit, jax.vmap, jax.value_and_grad, jax.Array = jnp, _jit, tvmap, value_and_grad, torch.Tensor
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.config_module = jax.config
    lax = types.ModuleType("jax.lax")
    lax.map = lambda f, xs, **_k: tvmap(f)(xs)
    jax.lax = lax
    jss = types.ModuleType("jax.scipy.stats")
    inv_sqrt_2pi = 1.0 / math.sqrt(2.0 * math.pi)
    jss.norm = types.SimpleNamespace(pdf=lambda u: inv_sqrt_2pi * torch.exp(-0.5 * u * u),
                                     cdf=lambda u: 0.5 * torch.erfc(-u / math.sqrt(2.0)))
    tfp_root = types.ModuleType("tensorflow_probability")
    tfp_root.__path__ = []
    tfp_sub = types.ModuleType("tensorflow_probability.substrates")
    tfp_sub.__path__ = []
    tfj = types.ModuleType("tensorflow_probability.substrates.jax")

    def t_log1mexp(x):  # TFP: log(1 - exp(-|x|))
        x = torch.abs(x)
        return torch.where(x < math.log(2.0), torch.log(-torch.expm1(-x)), torch.log1p(-torch.exp(-x)))
    tfj.math = types.SimpleNamespace(erfcx=torch.special.erfcx, log1mexp=t_log1mexp)
    tfp_root.substrates, tfp_sub.jax = tfp_sub, tfj
    rnd = types.ModuleType("jax.random")
    rnd.PRNGKey = lambda seed: seed
    jax.random = rnd
    jsp, jsl = types.ModuleType("jax.scipy"), types.ModuleType("jax.scipy.linalg")
    jsp.__path__ = []
    jsl.cho_solve, jsl.solve_triangular = cho_solve, solve_triangular
    jsp.linalg, jsp.stats = jsl, jss
    jax.scipy = jsp
    dist = types.ModuleType("numpyro.distributions")

    class Distribution:
        pass

    class Uniform(Distribution):
        def __init__(self, low=0.0, high=1.0):
            self.low, self.high = low, high

        def log_prob(self, x):
            return -math.log(self.high - self.low) + 0.0 * x

    class LogNormal(Distribution):  # textbook densities, differentiable (checked against scipy.stats by the NumPy-run values)
        def __init__(self, loc=0.0, scale=1.0):
            self.loc, self.scale = float(loc), float(scale)

        def log_prob(self, x):
            x = as_t(x)
            return -torch.log(x) - math.log(self.scale) - 0.5 * math.log(2.0 * math.pi) - (torch.log(x) - self.loc) ** 2 / (2.0 * self.scale ** 2)

    class HalfCauchy(Distribution):
        def __init__(self, scale=1.0):
            self.scale = float(scale)

        def log_prob(self, x):
            x = as_t(x)
            return math.log(2.0 / math.pi) - math.log(self.scale) - torch.log1p((x / self.scale) ** 2)
    dist.Distribution, dist.Uniform, dist.LogNormal, dist.HalfCauchy = Distribution, Uniform, LogNormal, HalfCauchy
    numpyro = types.ModuleType("numpyro")
    numpyro.__path__ = []
    numpyro.distributions = dist
    names = ("jax", "jax.numpy", "jax.random", "jax.lax", "jax.scipy", "jax.scipy.linalg", "jax.scipy.stats", "numpyro",
             "numpyro.distributions", "tensorflow_probability", "tensorflow_probability.substrates",
             "tensorflow_probability.substrates.jax")
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(dict(zip(names, (jax, jnp, rnd, lax, jsp, jsl, jss, numpyro, dist, tfp_root, tfp_sub, tfj))))
    try:
        pkg = types.ModuleType("BOBE_ad")
        pkg.__path__ = [os.path.join(REF, "BOBE")]
        utils = types.ModuleType("BOBE_ad.utils")
        utils.__path__ = [os.path.join(REF, "BOBE", "utils")]
        samplers = types.ModuleType("BOBE_ad.samplers")
        samplers.nested_sampling_Dy = samplers.sample_GP_NUTS = None
        sys.modules.update({"BOBE_ad": pkg, "BOBE_ad.utils": utils, "BOBE_ad.samplers": samplers})
        mods = {}
        for name, rel in (("BOBE_ad.utils.log", "utils/log.py"), ("BOBE_ad.utils.seed", "utils/seed.py"),
                          ("BOBE_ad.utils.core", "utils/core.py"), ("BOBE_ad.optim", "optim.py"), ("BOBE_ad.gp", "gp.py"),
                          ("BOBE_ad.acquisition", "acquisition.py")):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "BOBE", rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        for k, m in saved.items():
            if m is not None:
                sys.modules[k] = m
    return mods["BOBE_ad.gp"], jax, mods["BOBE_ad.acquisition"]


# ---- the cases ------------------------------------------------------------------------------------------------------
def _objective(X):
    return -0.5 * np.sum(((X - 0.45) / 0.2) ** 2, axis=1) + 0.3 * np.sin(7.0 * X[:, 0]) + 12.5


def _training_set(rng, n, d):
    X = rng.uniform(0.0, 1.0, size=(n, d))
    return X, _objective(X)


def generate():
    G, A, C = load_reference()
    G_ad, jax_ad, A_ad = load_reference_autodiff()
    rng = np.random.default_rng(20261018)
    out, state_vals, state_keys, state_meta = {}, {}, [], {}

    # kernels and distances --------------------------------------------------------------------------------------------
    xa, xb = rng.uniform(0, 1, (40, 3)), rng.uniform(0, 1, (7, 3))
    ls, kv, noise = np.array([0.31, 0.8, 1.7]), 1.7, 1e-6
    out.update(k_xa=xa, k_xb=xb, k_ls=ls, k_kv=kv, k_noise=noise,
               k_dist_sq=G.dist_sq(xa, xb),
               k_rbf_cross=G.rbf_kernel(xa, xb, ls, kv, noise, include_noise=False),
               k_matern_cross=G.matern_kernel(xa, xb, ls, kv, noise, include_noise=False),
               k_rbf_square=G.rbf_kernel(xa, xa, ls, kv, noise, include_noise=True),
               k_matern_square=G.matern_kernel(xa, xa, ls, kv, noise, include_noise=True),
               k_diag_noise=G.kernel_diag(xa, kv, noise, include_noise=True),
               k_diag_plain=G.kernel_diag(xa, kv, noise, include_noise=False))

    # log marginal likelihood of a given K ------------------------------------------------------------------------------
    X, y = _training_set(rng, 60, 3)
    ys = ((y - y.mean()) / y.std())[:, None]
    for kern in ("rbf", "matern"):
        K = getattr(G, kern + "_kernel")(X, X, np.array([0.4, 0.6, 0.9]), 1.3, 1e-6, include_noise=True)
        out[f"mll_{kern}_K"] = K
        out[f"mll_{kern}_value"] = np.float64(G.gp_mll(K, ys, 60))
    out.update(mll_X=X, mll_y_std=ys)

    # rank-1 Cholesky append ------------------------------------------------------------------------------------------
    Xc = rng.uniform(0, 1, (30, 2))
    Kc = G.matern_kernel(Xc, Xc, np.array([0.5, 0.7]), 1.0, 1e-6, include_noise=True)
    Lc = np.linalg.cholesky(Kc)
    xnew = rng.uniform(0, 1, (1, 2))
    kvec = G.matern_kernel(Xc, xnew, np.array([0.5, 0.7]), 1.0, 1e-6, include_noise=False).flatten()
    out.update(chol_L=Lc, chol_k=kvec, chol_kself=np.float64(1.0 + 1e-6),
               chol_new_L=np.asarray(G.fast_update_cholesky(Lc, kvec, 1.0 + 1e-6)))

    # the GP class ------------------------------------------------------------------------------------------------------
    # two well-conditioned cases (factor stored) and BASELINE config B's worst-conditioned shape (n = 500, d = 2 RBF with
    # lengthscale 0.3, noise 1e-8: cond(K) ~ 1e10; factor not stored -- 2 MB)
    for p, kern, n, d, m, lsg, kvg, noise_g, store_factor in (
            ("gp_rbf_", "rbf", 80, 4, 33, np.array([0.35, 0.6, 0.9, 1.4]), 1.3, 1e-6, True),
            ("gp_matern_", "matern", 80, 4, 33, np.array([0.35, 0.6, 0.9, 1.4]), 1.3, 1e-6, True),
            ("gpB_rbf_", "rbf", 500, 2, 40, np.array([0.3, 0.3]), 1.0, 1e-8, False)):
        X, y = _training_set(rng, n, d)
        gp = G.GP(X, y[:, None], noise=noise_g, kernel=kern, lengthscales=lsg, kernel_variance=kvg)
        Xq = rng.uniform(0, 1, (m, d))
        Xq[0] = X[5]  # a training point: variance at the noise floor
        out.update({p + "X": X, p + "y": y, p + "ls": lsg, p + "kv": np.float64(kvg), p + "noise": np.float64(noise_g), p + "Xq": Xq,
                    p + "y_mean": np.float64(gp.y_mean), p + "y_std": np.float64(gp.y_std),
                    p + "train_y": np.asarray(gp.train_y), p + "alphas": np.asarray(gp.alphas),
                    p + "cond_L": np.float64(np.linalg.cond(np.asarray(gp.cholesky))),
                    p + "logdet_half": np.float64(np.sum(np.log(np.diag(np.asarray(gp.cholesky))))),
                    p + "mean_batched": np.asarray(gp.predict_mean_batched(Xq)),
                    p + "var_batched": np.asarray(gp.predict_var_batched(Xq)),
                    p + "mean_single": np.float64(gp.predict_mean_single(Xq[3])),
                    p + "var_single": np.float64(gp.predict_var_single(Xq[3]))})
        if store_factor:
            out[p + "cholesky"] = np.asarray(gp.cholesky)
        ms, vs = gp.predict_batched(Xq)
        out.update({p + "std_mean_batched": np.asarray(ms).reshape(-1), p + "std_var_batched": np.asarray(vs).reshape(-1)})

        # neg_mll at several hyper-parameter rows (default priors are Uniform: a constant, evaluated by the stand-in) and
        # central differences of it (autodiff is not available; a sanity bound for the analytic gradient)
        ls_hi = 2.0 if store_factor else 0.45  # (the dense d = 2 shape turns numerically singular beyond that)
        lp = np.log(np.column_stack([rng.uniform(0.2, ls_hi, (5, d)), rng.uniform(0.5, 3.0, 5)]))
        vals = np.array([float(gp.neg_mll(r)) for r in lp])
        h = 1e-5
        fd = np.zeros_like(lp)
        for r in range(lp.shape[0]):
            for j in range(lp.shape[1]):
                e = np.zeros(lp.shape[1])
                e[j] = h
                fd[r, j] = (float(gp.neg_mll(lp[r] + e)) - float(gp.neg_mll(lp[r] - e))) / (2 * h)
        prior_const = float(gp.prior_func(np.exp(lp[0, :d]), np.exp(lp[0, d]), 1.0))
        out.update({p + "log_params": lp, p + "neg_mll": vals, p + "neg_mll_fd_grad": fd, p + "prior_const": np.float64(prior_const)})
        # ... and jax.value_and_grad(neg_mll) as BOBE/optim.py:307-309 builds it, run as reverse-mode autodiff through the
        # reference's own statements (torch-backed stand-in)
        import torch
        gp_ad = G_ad.GP(X, y[:, None], noise=noise_g, kernel=kern, lengthscales=torch.as_tensor(lsg), kernel_variance=kvg)
        vg = jax_ad.value_and_grad(gp_ad.neg_mll)
        ad = [vg(r) for r in lp]
        out.update({p + "neg_mll_ad": np.array([a[0] for a in ad]), p + "neg_mll_ad_grad": np.stack([a[1] for a in ad])})
        if store_factor:
            # GP.fit -> optimize_scipy (BOBE/gp.py:400-437, BOBE/optim.py:249-359): L-BFGS-B from four starting points on
            # the reference's own value_and_grad, best finite result
            lo, hi = np.asarray(gp_ad.hyperparam_bounds[0]), np.asarray(gp_ad.hyperparam_bounds[1])
            x0 = np.clip(np.log(np.column_stack([rng.uniform(0.1, 2.5, (4, d)), rng.uniform(0.3, 5.0, 4)])), lo, hi)
            res = gp_ad.fit(x0, maxiter=80)
            out.update({p + "fit_x0": x0, p + "fit_mll": np.float64(res["mll"]), p + "fit_params": np.asarray(res["params"])})

        # fantasy variance and the two integrated acquisitions
        mc = rng.uniform(0, 1, (50, d))
        cand = rng.uniform(0, 1, (3, d))
        k_train_mc = gp.kernel(gp.train_x, mc, gp.lengthscales, gp.kernel_variance, noise=gp.noise, include_noise=False)
        fv = np.stack([np.asarray(gp.fantasy_var(c, mc, k_train_mc)) for c in cand])
        out.update({p + "mc": mc, p + "cand": cand, p + "fantasy_var": fv,
                    p + "wipv": np.array([float(A.WIPV().fun(c, gp, mc_points=mc, k_train_mc=k_train_mc)) for c in cand]),
                    p + "wipstd": np.array([float(A.WIPStd().fun(c, gp, mc_points=mc, k_train_mc=k_train_mc)) for c in cand])})

        # EI / LogEI (negated, as the optimiser sees them) on the standardised scale
        best_y, zeta = float(np.max(np.asarray(gp.train_y))), 0.01
        xe = np.vstack([Xq[:8], X[5][None, :]])
        out.update({p + "ei_x": xe, p + "ei_best_y": np.float64(best_y), p + "ei_zeta": np.float64(zeta),
                    p + "ei": np.array([float(A.EI().fun(x, gp, best_y, zeta)) for x in xe]),
                    p + "logei": np.array([float(A.LogEI().fun(x, gp, best_y, zeta)) for x in xe])})

        # jax.value_and_grad of the acquisition functions in the candidate point (what BOBE/optim.py differentiates when it
        # polishes EI / LogEI / WIPV / WIPStd): reverse mode through predict_single / fantasy_var of the reference
        mc_t = torch.as_tensor(mc)
        k_train_mc_t = gp_ad.kernel(gp_ad.train_x, mc_t, gp_ad.lengthscales, gp_ad.kernel_variance, noise=gp_ad.noise,
                                    include_noise=False)
        xg = np.vstack([Xq[1:6], cand])  # (query points away from the training set, and the WIPV candidates)
        acq_grads = {}
        for name, obj, args, kwargs in (("ei", A_ad.EI(), (gp_ad, best_y, zeta), {}), ("logei", A_ad.LogEI(), (gp_ad, best_y, zeta), {}),
                                        ("wipv", A_ad.WIPV(), (gp_ad,), {"mc_points": mc_t, "k_train_mc": k_train_mc_t}),
                                        ("wipstd", A_ad.WIPStd(), (gp_ad,), {"mc_points": mc_t, "k_train_mc": k_train_mc_t})):
            vg_acq = jax_ad.value_and_grad(obj.fun)
            res_acq = [vg_acq(x, *args, **kwargs) for x in xg]
            acq_grads[p + name + "_ad_value"] = np.array([r[0] for r in res_acq])
            acq_grads[p + name + "_ad_grad"] = np.stack([r[1] for r in res_acq])
        # ... and of the standardised posterior mean / variance themselves (BOBE/gp.py:476-489)
        for name, fn in (("pmean", lambda x: jax_ad.numpy.reshape(gp_ad.predict_single(x)[0], ())),
                         ("pvar", lambda x: jax_ad.numpy.reshape(gp_ad.predict_single(x)[1], ()))):
            res_p = [jax_ad.value_and_grad(fn)(x) for x in xg]
            acq_grads[p + name + "_ad_value"] = np.array([r[0] for r in res_p])
            acq_grads[p + name + "_ad_grad"] = np.stack([r[1] for r in res_p])
        out.update(acq_grads)
        out[p + "acq_grad_x"] = xg
        if store_factor:
            # the acquisition optimisation flows themselves, with a seeded generator (BOBE/acquisition.py:255-291,350-412 ->
            # BOBE/optim.py:249-359): starting points from the rng, L-BFGS-B on jax.value_and_grad of the reference's fun
            for name, obj, kw_acq in (("ei", A_ad.EI(), {"zeta": zeta, "best_y": best_y}), ("logei", A_ad.LogEI(), {"zeta": zeta, "best_y": best_y})):
                pt, val = obj.get_next_point(gp_ad, kw_acq, maxiter=100, n_restarts=6, verbose=False, rng=np.random.default_rng(7))
                out.update({p + "flow_" + name + "_x": np.asarray(pt), p + "flow_" + name + "_val": np.float64(val)})
            mc_samples = {"x": mc}
            for name, obj in (("wipv", A_ad.WIPV()), ("wipstd", A_ad.WIPStd())):
                pt, val = obj.get_next_point(gp_ad, {"mc_samples": {"x": torch.as_tensor(mc)}, "mc_points_size": 32}, maxiter=60,
                                             n_restarts=1, verbose=False, rng=np.random.default_rng(11))
                out.update({p + "flow_" + name + "_x": np.asarray(pt), p + "flow_" + name + "_val": np.float64(val)})
        if p == "gp_matern_":  # the state dictionary the reference saves / loads / copies through (BOBE/gp.py:586-636)
            st = gp.state_dict()
            state_keys = sorted(st.keys())
            for k, val in st.items():
                if isinstance(val, np.ndarray) or isinstance(val, (int, float, bool)):
                    state_vals[k] = np.asarray(val, dtype=np.float64)
            state_meta = {k: (val if isinstance(val, (str, dict, list, bool, int, float)) or val is None else None)
                          for k, val in st.items() if not isinstance(val, np.ndarray)}
        # update(): two new points and one duplicate of a training point
        new_x = np.vstack([rng.uniform(0, 1, (2, d)), X[7][None, :]])
        new_y = np.concatenate([_objective(new_x[:2]) + np.array([0.01, -0.02]), [y[7]]])[:, None]
        gp.update(new_x, new_y)
        out.update({p + "upd_new_x": new_x, p + "upd_new_y": new_y, p + "upd_train_x": np.asarray(gp.train_x),
                    p + "upd_y_mean": np.float64(gp.y_mean), p + "upd_y_std": np.float64(gp.y_std),
                    p + "upd_alphas": np.asarray(gp.alphas),
                    p + "upd_cond_L": np.float64(np.linalg.cond(np.asarray(gp.cholesky))),
                    p + "upd_mean_batched": np.asarray(gp.predict_mean_batched(Xq[:6]))})
        if store_factor:
            out[p + "upd_cholesky"] = np.asarray(gp.cholesky)

    # the HEADLINE shape (BASELINE config H: n = 2000, d = 16, Matern-5/2 ARD): inputs are the seeded synthetic set of
    # oracle.gp_oracle (regenerated by the tests, not stored); outputs of the reference's own code only
    sys.path.insert(0, ROOT)
    from oracle import gp_oracle as O
    nH, dH = 2000, 16
    XH, yH = O.synthetic_training_set(nH, dH)
    lsH = np.ones(dH)
    gpH = G.GP(XH, np.asarray(yH).reshape(-1, 1), noise=1e-8, kernel="matern", lengthscales=lsH, kernel_variance=1.0)
    XqH = O.synthetic_queries(48, dH, seed=21)
    msH, vsH = gpH.predict_batched(XqH)
    mcH, candH = O.synthetic_queries(64, dH, seed=22), O.synthetic_queries(2, dH, seed=23)
    ktmH = gpH.kernel(gpH.train_x, mcH, gpH.lengthscales, gpH.kernel_variance, noise=gpH.noise, include_noise=False)
    lpH = np.log(np.column_stack([rng.uniform(0.6, 1.8, (3, dH)), rng.uniform(0.6, 2.0, 3)]))
    import torch
    gpH_ad = G_ad.GP(XH, np.asarray(yH).reshape(-1, 1), noise=1e-8, kernel="matern", lengthscales=torch.as_tensor(lsH), kernel_variance=1.0)
    adH = [jax_ad.value_and_grad(gpH_ad.neg_mll)(r) for r in lpH]
    out.update(gpH_n=nH, gpH_d=dH, gpH_y_mean=np.float64(gpH.y_mean), gpH_y_std=np.float64(gpH.y_std),
               gpH_cond_L=np.float64(np.linalg.cond(np.asarray(gpH.cholesky))),
               gpH_logdet_half=np.float64(np.sum(np.log(np.diag(np.asarray(gpH.cholesky))))),
               gpH_alphas=np.asarray(gpH.alphas), gpH_mean_batched=np.asarray(gpH.predict_mean_batched(XqH)),
               gpH_var_batched=np.asarray(gpH.predict_var_batched(XqH)), gpH_std_mean_batched=np.asarray(msH).reshape(-1),
               gpH_std_var_batched=np.asarray(vsH).reshape(-1),
               gpH_fantasy_var=np.stack([np.asarray(gpH.fantasy_var(c, mcH, ktmH)) for c in candH]),
               gpH_log_params=lpH, gpH_neg_mll=np.array([float(gpH.neg_mll(r)) for r in lpH]),
               gpH_neg_mll_ad=np.array([a[0] for a in adH]), gpH_neg_mll_ad_grad=np.stack([a[1] for a in adH]))

    # BASELINE config D (n = 1500, d = 27, RBF, lengthscale 2) and config E (n = 4000, d = 12, RBF: the WIPV shape), same
    # convention: seeded inputs regenerated by the tests, the reference's outputs stored
    for tag, nS, dS, ellS in (("gpD_", 1500, 27, 2.0), ("gpE_", 4000, 12, 1.0)):
        XS, yS = O.synthetic_training_set(nS, dS)
        gpS = G.GP(XS, np.asarray(yS).reshape(-1, 1), noise=1e-8, kernel="rbf", lengthscales=np.full(dS, ellS), kernel_variance=1.0)
        XqS = O.synthetic_queries(32, dS, seed=31)
        mcS, candS = O.synthetic_queries(48, dS, seed=32), O.synthetic_queries(2, dS, seed=33)
        ktmS = gpS.kernel(gpS.train_x, mcS, gpS.lengthscales, gpS.kernel_variance, noise=gpS.noise, include_noise=False)
        fvS = np.stack([np.asarray(gpS.fantasy_var(c, mcS, ktmS)) for c in candS])
        out.update({tag + "n": nS, tag + "d": dS, tag + "ell": ellS, tag + "y_std": np.float64(gpS.y_std),
                    tag + "cond_L": np.float64(np.linalg.cond(np.asarray(gpS.cholesky))),
                    tag + "logdet_half": np.float64(np.sum(np.log(np.diag(np.asarray(gpS.cholesky))))),
                    tag + "mean_batched": np.asarray(gpS.predict_mean_batched(XqS)),
                    tag + "var_batched": np.asarray(gpS.predict_var_batched(XqS)),
                    tag + "fantasy_var": fvS, tag + "wipv": fvS.mean(axis=1), tag + "wipstd": np.sqrt(fvS).mean(axis=1)})

    # BASELINE config A (n = 100, d = 2, RBF with lengthscale 0.3: cond(K) ~ 3e9): WIPV / WIPStd with the 512 MC points as
    # their own candidates, exactly the sweep of BOBE/acquisition.py:385-397 (lax.map of fun over mc_points)
    XA, yA = O.synthetic_training_set(100, 2)
    gpA = G.GP(XA, np.asarray(yA).reshape(-1, 1), noise=1e-8, kernel="rbf", lengthscales=np.full(2, 0.3), kernel_variance=1.0)
    mcA = O.synthetic_queries(512, 2, seed=5)
    ktmA = gpA.kernel(gpA.train_x, mcA, gpA.lengthscales, gpA.kernel_variance, noise=gpA.noise, include_noise=False)
    out.update(gpA_y_std=np.float64(gpA.y_std), gpA_cond_L=np.float64(np.linalg.cond(np.asarray(gpA.cholesky))),
               gpA_wipv_self=np.array([float(A.WIPV().fun(x, gpA, mc_points=mcA, k_train_mc=ktmA)) for x in mcA]),
               gpA_wipstd_self=np.array([float(A.WIPStd().fun(x, gpA, mc_points=mcA, k_train_mc=ktmA)) for x in mcA]),
               gpA_mean_batched=np.asarray(gpA.predict_mean_batched(mcA[:64])), gpA_var_batched=np.asarray(gpA.predict_var_batched(mcA[:64])))

    # priors: DSLP lengthscales + LogNormal kernel variance, and SAAS (adds tausq as a hyper-parameter)
    import torch
    n, d = 60, 3
    X, y = _training_set(rng, n, d)
    for tag, kwargs in (("prior_dslp_", dict(lengthscale_prior="DSLP", kernel_variance_prior={"name": "LogNormal", "loc": 0.0, "scale": 1.0})),
                        ("prior_saas_", dict(lengthscale_prior="SAAS", tausq=0.7)),
                        ("prior_fixedkv_", dict(lengthscale_prior="DSLP", kernel_variance_prior="fixed"))):
        gpp = G.GP(X, y[:, None], noise=1e-6, kernel="matern", lengthscales=np.array([0.5, 0.8, 1.1]), kernel_variance=1.4, **kwargs)
        gpt = G_ad.GP(X, y[:, None], noise=1e-6, kernel="matern", lengthscales=torch.as_tensor(np.array([0.5, 0.8, 1.1])),
                      kernel_variance=1.4, **kwargs)
        P = int(gpp.num_hyperparams)
        lp = np.log(rng.uniform(0.3, 2.0, (4, P)))
        vg = jax_ad.value_and_grad(gpt.neg_mll)
        ad = [vg(r) for r in lp]
        out.update({tag + "X": X, tag + "y": y, tag + "log_params": lp, tag + "num_hyperparams": P,
                    tag + "neg_mll": np.array([float(gpp.neg_mll(r)) for r in lp]),
                    tag + "prior": np.array([float(np.sum(gpp.prior_func(*gpp._parse_hyperparams(r)))) for r in lp]),
                    tag + "neg_mll_ad": np.array([a[0] for a in ad]), tag + "neg_mll_ad_grad": np.stack([a[1] for a in ad]),
                    tag + "hyperparam_bounds": np.asarray(gpp.hyperparam_bounds)})

    # GPwithClassifier with the SVM mask (BOBE/clf_gp.py, BOBE/clf.py:36-83,188-214): a target with a deep infeasible region
    n, d, m = 120, 3, 60
    X = rng.uniform(0, 1, (n, d))
    y = -60.0 * np.sum((X - 0.5) ** 2, axis=1) + 0.5 * np.sin(5.0 * X[:, 1])
    cgp = C.GPwithClassifier(X, y[:, None], clf_type="svm", clf_use_size=10, clf_threshold=8.0, gp_threshold=16.0, noise=1e-6,
                             kernel="rbf", lengthscales=np.array([0.4, 0.5, 0.6]), kernel_variance=1.2,
                             lengthscale_prior={"name": "Uniform", "low": 0.01, "high": 5.0})
    Xq = rng.uniform(0, 1, (m, d))
    assert cgp.use_clf and cgp._clf_predict_func is not None
    cm, cv = cgp.predict_batched(Xq)
    mask = np.array([float(cgp._clf_predict_func(x)) for x in Xq])
    assert 0 < mask.sum() < m  # both sides of the mask are exercised
    out.update(clf_X=X, clf_y=y, clf_Xq=Xq, clf_ls=np.array([0.4, 0.5, 0.6]), clf_kv=1.2, clf_noise=1e-6,
               clf_threshold=8.0, clf_gp_threshold=16.0, clf_minus_inf=cgp.minus_inf,
               clf_gp_train_x=np.asarray(cgp.train_x), clf_y_mean=np.float64(cgp.y_mean), clf_y_std=np.float64(cgp.y_std),
               clf_support_vectors=np.asarray(cgp.clf_params["support_vectors"]), clf_dual_coef=np.asarray(cgp.clf_params["dual_coef"]),
               clf_intercept=cgp.clf_params["intercept"], clf_gamma=cgp.clf_params["gamma_eff"],
               clf_decision=np.array([float(C.svm_predict(x, cgp.clf_params["support_vectors"], cgp.clf_params["dual_coef"],
                                                          cgp.clf_params["intercept"], cgp.clf_params["gamma_eff"])) for x in Xq]),
               clf_mask=mask, clf_mean_batched=np.asarray(cgp.predict_mean_batched(Xq)),
               clf_var_batched=np.asarray(cgp.predict_var_batched(Xq)),
               clf_std_mean_batched=np.asarray(cm).reshape(-1), clf_std_var_batched=np.asarray(cv).reshape(-1))

    # the LogEI helper over its three branches (u > -1, the asymptotic branch, and u < -1e6)
    u = np.concatenate([np.linspace(-40.0, 5.0, 91), -np.logspace(2, 7.5, 12)])
    out.update(logei_u=u, logei_helper=np.asarray(A._log_ei_helper(u)), ei_helper=np.asarray(A._ei_helper(u)))

    out = {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}
    out.update({"state_" + k: v for k, v in state_vals.items()})
    import inspect
    import json

    # the public surface of the path: every public function / method with its parameter names and defaults, for the
    # drop-in check of tests/test_host_logic.py (same names, same order, same defaults)
    def sig(fn):
        try:
            ps = inspect.signature(fn).parameters.values()
        except (TypeError, ValueError):
            return None
        return [[q.name, None if q.default is inspect.Parameter.empty else repr(q.default)] for q in ps]

    surface = {"functions": {}, "classes": {}}
    opt_mod = sys.modules["BOBE.optim"]
    for mod, names in ((G, ("dist_sq", "kernel_diag", "rbf_kernel", "matern_kernel", "gp_mll", "fast_update_cholesky",
                            "saas_prior_logprob", "make_distribution")),
                       (opt_mod, ("optimize_scipy", "optimize_optax", "optimize_optax_vmap"))):
        for name in names:
            surface["functions"][name] = sig(getattr(mod, name))
    for mod, names in ((G, ("GP",)), (C, ("GPwithClassifier",)), (A, ("AcquisitionFunction", "EI", "LogEI", "WIPV", "WIPStd"))):
        for name in names:
            cls = getattr(mod, name)
            members = {}
            for mname, member in inspect.getmembers(cls):
                if mname.startswith("_") and mname != "__init__":
                    continue
                if isinstance(inspect.getattr_static(cls, mname), property):
                    members[mname] = "property"
                elif callable(member):
                    members[mname] = sig(member)
            surface["classes"][name] = members
    with open(os.path.join(ROOT, "tests", "golden", "reference_public_surface.json"), "w") as f:
        json.dump(surface, f, indent=1, sort_keys=True)
    out["state_keys_json"] = np.array(json.dumps(state_keys))
    out["state_meta_json"] = np.array(json.dumps(state_meta))
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    generate()
