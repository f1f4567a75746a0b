"""CPU oracle for the GP surrogate hot path of Ameek94/BOBE -- TEST INFRASTRUCTURE ONLY.

This is a NumPy/SciPy float64 restatement of the arithmetic in the reference's ``BOBE/gp.py`` and of
the acquisition arithmetic in ``BOBE/acquisition.py``.  It exists to check the CUDA path; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``bobe_b200/`` imports it.

PINNING.  The reference is pure Python/JAX; JAX, numpyro and tensorflow-probability are not installable in this image
(no wheel, no network), so ``import BOBE`` fails and the reference's own tests hold no golden vectors for this path
(SURVEY.md 8c).  Two things stand in for a JAX run:

1. THE REFERENCE'S OWN SOURCE, EXECUTED HERE (round 2): ``oracle/gen_reference_vectors.py`` loads the unmodified files
   ``BOBE/gp.py`` and ``BOBE/acquisition.py`` from /root/reference under a NumPy/SciPy stand-in for the slice of the jax API
   they touch (jax.numpy -> numpy, jax.scipy.linalg -> scipy.linalg, jit -> identity, vmap / lax.map -> loops) and stores
   what the reference's code computes -- kernels, gp_mll, fast_update_cholesky, the GP constructor's standardisation /
   Cholesky / alphas, every predict variant, neg_mll, update() with a duplicate, fantasy_var, WIPV / WIPStd / EI / LogEI
   values, at two well-conditioned shapes and at BASELINE config B's worst-conditioned one -- in
   ``tests/golden/reference_source_vectors.npz``.  ``tests/test_oracle.py`` holds this restatement to those vectors (the
   factor and the kernels agree bit for bit, everything else to ~1e-14 x cond), and ``tests/test_gpu_parity.py`` holds the
   CUDA path to them.  The log-ML GRADIENT is pinned the same way: gp.py is loaded a second time with jax.numpy -> torch
   float64 tensors, where ``jax.value_and_grad(gp.neg_mll)`` (BOBE/optim.py:307-309) runs as reverse-mode autodiff through
   the reference's own statements; the analytic gradient below agrees with it to 1e-12 .. 3e-10 (9e-8 at cond(K) ~ 1e10).
   The same load runs the reference's GP.fit -> optimize_scipy, differentiates predict_single / EI / LogEI / WIPV / WIPStd in
   the query point, and evaluates the DSLP / SAAS / fixed-kv prior compositions (scipy.stats densities standing in for
   numpyro's).  NOT pinned by this route: XLA's own floating-point behaviour (rounding-level) and numpyro's density code.
2. MATHEMATICS AND THIRD-PARTY CODE (round 1, tests/test_oracle.py): closed forms at n=1,2, interpolation / noise-level
   identities, the gradient three ways (analytic, torch autograd through ``torch.linalg.cholesky`` = the reverse-mode
   construction JAX uses, central differences), ``fantasy_var`` == ``predict_var`` of an actually-updated GP, an mpmath
   60-digit re-evaluation (``oracle/truth_mp.py``), scikit-learn's ``GaussianProcessRegressor`` (posterior mean / variance,
   log marginal likelihood and its gradient, both kernels) and ``SVC.decision_function`` (the classifier mask).

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla
import scipy.special as ssp

SAFE_NOISE_FLOOR = 1e-12  # BOBE/gp.py:16
SQRT2 = math.sqrt(2.0)  # BOBE/gp.py:18
SQRT3 = math.sqrt(3.0)  # BOBE/gp.py:19
SQRT5 = math.sqrt(5.0)  # BOBE/gp.py:20
LOG_2PI = math.log(2.0 * math.pi)
DIST_SQ_TEMP_ELEMS = 2**18  # size cap of the (rows, n2, d) temporary in dist_sq; arithmetic per entry unchanged


# ----------------------------------------------------------------------------------------------
# kernels  (BOBE/gp.py:80-168)
# ----------------------------------------------------------------------------------------------
def dist_sq(x, y):
    """BOBE/gp.py:80-96 -- sum_k (x_ik - y_jk)^2 by direct differences."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    out = np.empty((x.shape[0], y.shape[0]))
    # row-chunked only to bound the (n1,n2,d) temporary; the arithmetic per entry is unchanged
    step = max(1, int(DIST_SQ_TEMP_ELEMS // max(1, y.shape[0] * x.shape[1])))
    for s in range(0, x.shape[0], step):
        diff = x[s:s + step, None, :] - y[None, :, :]
        out[s:s + step] = np.sum(np.square(diff), axis=-1)
    return out


def kernel_diag(x, kernel_variance, noise, include_noise=True):
    """BOBE/gp.py:98-122."""
    diag = kernel_variance * np.ones(np.asarray(x).shape[0])
    if include_noise:
        diag = diag + noise
    return diag


def rbf_kernel(xa, xb, lengthscales, kernel_variance, noise, include_noise=True):
    """BOBE/gp.py:124-154 -- inputs are DIVIDED by the lengthscales (:149)."""
    ls = np.asarray(lengthscales, dtype=np.float64)
    sq = dist_sq(np.asarray(xa) / ls, np.asarray(xb) / ls)
    k = kernel_variance * np.exp(-0.5 * sq)
    if include_noise:
        k = k + noise * np.eye(k.shape[0])
    return k


def matern_kernel(xa, xb, lengthscales, kernel_variance, noise, include_noise=True):
    """BOBE/gp.py:156-168 -- Matern-5/2 with the 1e-30 clamp under the sqrt (:162)."""
    ls = np.asarray(lengthscales, dtype=np.float64)
    dsq = dist_sq(np.asarray(xa) / ls, np.asarray(xb) / ls)
    d = np.sqrt(np.where(dsq < 1e-30, 1e-30, dsq))
    e = np.exp(-SQRT5 * d)
    poly = 1.0 + d * (SQRT5 + d * 5.0 / 3.0)
    k = kernel_variance * poly * e
    if include_noise:
        k = k + noise * np.eye(k.shape[0])
    return k


def get_kernel(name):
    """BOBE/gp.py:251-252 -- anything that is not "rbf" is Matern."""
    return rbf_kernel if name == "rbf" else matern_kernel


# ----------------------------------------------------------------------------------------------
# factor, marginal likelihood, rank-1 append  (BOBE/gp.py:170-197)
# ----------------------------------------------------------------------------------------------
def cholesky_nan(k):
    """jnp.linalg.cholesky semantics (BOBE/gp.py:175,259,549): lower factor, NaN (no raise) if not PD."""
    try:
        return np.linalg.cholesky(k)
    except np.linalg.LinAlgError:
        return np.full_like(k, np.nan)


def cho_solve_lower(L, b):
    """jax.scipy.linalg.cho_solve((L, True), b) (BOBE/gp.py:176,260,550)."""
    if not np.all(np.isfinite(L)):
        return np.full_like(np.asarray(b, dtype=np.float64), np.nan)
    z = sla.solve_triangular(L, b, lower=True, check_finite=False)
    return sla.solve_triangular(L, z, lower=True, trans="T", check_finite=False)


def gp_mll(k, train_y, num_points):
    """BOBE/gp.py:170-178 -- log marginal likelihood (the docstring there says "negative"; the code is not)."""
    L = cholesky_nan(k)
    alpha = cho_solve_lower(L, train_y)
    return float(-0.5 * (train_y.T @ alpha).item() - np.sum(np.log(np.diag(L))) - 0.5 * num_points * LOG_2PI)


def fast_update_cholesky(L, k, k_self):
    """BOBE/gp.py:181-197 -- append one row/column to a lower Cholesky factor."""
    n = L.shape[0]
    v = sla.solve_triangular(L, k, lower=True, check_finite=False) if n else np.zeros(0)
    with np.errstate(invalid="ignore"):
        diag = np.sqrt(k_self - np.dot(v, v))
    new_L = np.zeros((n + 1, n + 1))
    new_L[:n, :n] = L
    new_L[n, :n] = v
    new_L[n, n] = diag
    return new_L


# ----------------------------------------------------------------------------------------------
# priors (numpyro.distributions 0.15.x log_prob restated; BOBE/gp.py:27-78,309-366)
# ----------------------------------------------------------------------------------------------
def uniform_logprob(z, low, high):
    """numpyro Uniform.log_prob: constant -log(high-low) (no support check without validate_args)."""
    return -math.log(high - low) * np.ones_like(np.asarray(z, dtype=np.float64))


def lognormal_logprob(z, loc, scale):
    """numpyro LogNormal = exp-transformed Normal: N(log z; loc, scale) - log z."""
    lz = np.log(z)
    return -0.5 * ((lz - loc) / scale) ** 2 - math.log(scale * math.sqrt(2 * math.pi)) - lz


def halfcauchy_logprob(z, scale):
    """numpyro HalfCauchy: log 2 + Cauchy(0, scale).log_prob(z)."""
    return math.log(2.0) - math.log(math.pi) - math.log(scale) - np.log1p((np.asarray(z) / scale) ** 2)


def saas_prior_logprob(lengthscales, kernel_variance, tausq):
    """BOBE/gp.py:56-78 (no Jacobian terms; tausq enters the prior only)."""
    lp = lognormal_logprob(kernel_variance, 0.0, 1.0)
    lp = lp + halfcauchy_logprob(tausq, 0.1)
    inv_ls_sq = 1.0 / (tausq * np.asarray(lengthscales) ** 2)
    lp = lp + np.sum(halfcauchy_logprob(inv_ls_sq, 1.0))
    return float(lp)


def standardise(train_y):
    """BOBE/gp.py:296-306 -- population std, std==0 -> 1."""
    train_y = np.asarray(train_y, dtype=np.float64)
    y_mean = float(np.mean(train_y)) if train_y.size > 0 else 0.0
    y_std = float(np.std(train_y)) if train_y.size > 0 else 1.0
    if y_std == 0:
        y_std = 1.0
    return (train_y - y_mean) / y_std, y_mean, y_std


class OracleGP:
    """Restatement of ``class GP`` (BOBE/gp.py:199-772) for the arithmetic the CUDA path replaces."""

    def __init__(self, train_x, train_y, noise=1e-8, kernel="rbf", lengthscales=None, kernel_variance=None,
                 kernel_variance_bounds=(1e-4, 1e8), lengthscale_bounds=(0.01, 5), kernel_variance_prior=None,
                 lengthscale_prior=None, tausq=None, tausq_bounds=(1e-4, 1e4)):
        train_x = np.asarray(train_x, dtype=np.float64)
        train_y = np.asarray(train_y, dtype=np.float64)
        if train_x.shape[0] != train_y.shape[0]:  # BOBE/gp.py:286-291
            raise ValueError("train_x and train_y must have the same number of points")
        if train_y.ndim != 2:
            train_y = train_y.reshape(-1, 1)
        if train_x.ndim != 2:
            raise ValueError("train_x must be 2D")
        self.ndim = train_x.shape[1]
        self.train_x = train_x
        self.train_y, self.y_mean, self.y_std = standardise(train_y)
        self.kernel_name = kernel if kernel == "rbf" else "matern"
        self.kernel = get_kernel(kernel)
        self.lengthscales = np.ones(self.ndim) if lengthscales is None else np.asarray(lengthscales, dtype=np.float64)
        self.kernel_variance = 1.0 if kernel_variance is None else float(kernel_variance)
        self.noise = float(noise)
        self.lengthscale_bounds = list(lengthscale_bounds)
        self.kernel_variance_bounds = list(kernel_variance_bounds)
        self.tausq = 1.0 if tausq is None else float(tausq)
        self.tausq_bounds = list(tausq_bounds)
        self.kernel_variance_prior_spec = kernel_variance_prior  # BOBE/gp.py:309-320
        if self.kernel_variance_prior_spec is None:
            self.kernel_variance_prior_spec = {"name": "Uniform", "low": self.kernel_variance_bounds[0],
                                               "high": self.kernel_variance_bounds[1]}
        self.fixed_kernel_variance = self.kernel_variance_prior_spec == "fixed"
        self.lengthscale_prior_spec = lengthscale_prior  # BOBE/gp.py:322-337
        if self.lengthscale_prior_spec is None:
            self.lengthscale_prior_spec = {"name": "Uniform", "low": self.lengthscale_bounds[0],
                                           "high": self.lengthscale_bounds[1]}
        bounds = [self.lengthscale_bounds] * self.ndim  # BOBE/gp.py:339-354
        if not self.fixed_kernel_variance:
            bounds.append(self.kernel_variance_bounds)
        if self.lengthscale_prior_spec == "SAAS":
            bounds.append(self.tausq_bounds)
        self.hyperparam_bounds = np.log(np.array(bounds, dtype=np.float64).T)
        self.num_hyperparams = self.hyperparam_bounds.shape[1]
        self.recompute_cholesky()

    # -- state ---------------------------------------------------------------------------------
    def recompute_cholesky(self):
        """BOBE/gp.py:544-550."""
        K = self.kernel(self.train_x, self.train_x, self.lengthscales, self.kernel_variance, self.noise, True)
        self.cholesky = cholesky_nan(K)
        self.alphas = cho_solve_lower(self.cholesky, self.train_y)

    @property
    def npoints(self):
        return self.train_x.shape[0]

    def get_hyperparams(self):
        """BOBE/gp.py:756-762."""
        hp = self.lengthscales
        if not self.fixed_kernel_variance:
            hp = np.hstack([hp, self.kernel_variance])
        if self.lengthscale_prior_spec == "SAAS":
            hp = np.hstack([hp, self.tausq])
        return hp

    def _parse_hyperparams(self, log_params):
        """BOBE/gp.py:368-383."""
        hp = np.exp(np.asarray(log_params, dtype=np.float64))
        ls = hp[:self.ndim]
        saas = self.lengthscale_prior_spec == "SAAS"
        if self.fixed_kernel_variance:
            kv = self.kernel_variance
            tausq = hp[self.ndim] if (saas and len(hp) > self.ndim) else self.tausq
        else:
            kv = hp[self.ndim]
            tausq = hp[self.ndim + 1] if len(hp) > self.ndim + 1 else self.tausq
        return ls, kv, tausq

    def update_hyperparams(self, log_params):
        """BOBE/gp.py:439-448."""
        ls, kv, tausq = self._parse_hyperparams(log_params)
        self.lengthscales = ls
        if not self.fixed_kernel_variance:
            self.kernel_variance = float(kv)
        self.tausq = float(tausq)
        self.recompute_cholesky()

    # -- priors --------------------------------------------------------------------------------
    def _dist_logprob_and_dlog(self, spec, z):
        """log p(z) and d log p / d log z for a numpyro-style spec dict (BOBE/gp.py:27-54)."""
        name = spec["name"]
        z = np.asarray(z, dtype=np.float64)
        if name == "Uniform":
            return uniform_logprob(z, spec["low"], spec["high"]), np.zeros_like(z)
        if name == "LogNormal":
            loc, scale = spec.get("loc", 0.0), spec.get("scale", 1.0)
            return lognormal_logprob(z, loc, scale), -(np.log(z) - loc) / scale**2 - 1.0
        if name == "HalfCauchy":
            s = spec.get("scale", 1.0)
            u2 = (z / s) ** 2
            return halfcauchy_logprob(z, s), -2.0 * u2 / (1.0 + u2)
        raise ValueError(f"Distribution {name} not found in numpyro.distributions.")

    def log_prior_and_grad(self, log_params):
        """BOBE/gp.py:357-366 + 56-78; gradient w.r.t. the log-parameters (same layout as log_params)."""
        ls, kv, tausq = self._parse_hyperparams(log_params)
        g = np.zeros(self.num_hyperparams)
        d = self.ndim
        if self.lengthscale_prior_spec == "SAAS":
            lp = saas_prior_logprob(ls, kv, tausq)
            u = 1.0 / (tausq * ls**2)
            w = u * u / (1.0 + u * u)
            g[:d] = 4.0 * w  # d/dlog(l_j) of -log(1+u_j^2), du/dlog l = -2u
            idx = d
            if not self.fixed_kernel_variance:
                g[idx] = -math.log(kv) - 1.0
                idx += 1
            if idx < self.num_hyperparams:
                t2 = (tausq / 0.1) ** 2
                g[idx] = -2.0 * t2 / (1.0 + t2) + 2.0 * np.sum(w)
            return float(lp), g
        lp = 0.0
        if not self.fixed_kernel_variance:
            l_kv, g_kv = self._dist_logprob_and_dlog(self.kernel_variance_prior_spec, kv)
            lp += float(l_kv)
            g[d] = float(g_kv)
        if self.lengthscale_prior_spec == "DSLP":
            spec = {"name": "LogNormal", "loc": SQRT2 + 0.5 * math.log(self.ndim), "scale": SQRT3}
        else:
            spec = self.lengthscale_prior_spec
        l_ls, g_ls = self._dist_logprob_and_dlog(spec, ls)
        lp += float(np.sum(l_ls))
        g[:d] = g_ls
        return lp, g

    # -- marginal likelihood -------------------------------------------------------------------
    def neg_mll(self, log_params):
        """BOBE/gp.py:385-398."""
        ls, kv, tausq = self._parse_hyperparams(log_params)
        K = self.kernel(self.train_x, self.train_x, ls, kv, self.noise, True)
        mll = gp_mll(K, self.train_y, self.train_y.shape[0])
        mll += self.log_prior_and_grad(log_params)[0]
        return -mll

    def neg_mll_and_grad(self, log_params):
        """value_and_grad(neg_mll) as the optimisers call it (BOBE/optim.py:118,211,309).

        Analytic: d log p / d theta = 1/2 sum_ik W_ik dK_ik/dtheta with W = alpha alpha^T - K^-1
        (SURVEY.md appendix A).  Cross-checked against torch autograd in tests/test_oracle.py.
        """
        ls, kv, tausq = self._parse_hyperparams(log_params)
        n, d = self.train_x.shape
        xs = self.train_x / ls
        q = dist_sq(xs, xs)
        if self.kernel_name == "rbf":
            K0 = kv * np.exp(-0.5 * q)
            G = K0  # dK/dlog l_j = G * s_j
        else:
            clamped = q < 1e-30
            r = np.sqrt(np.where(clamped, 1e-30, q))
            e = np.exp(-SQRT5 * r)
            K0 = kv * (1.0 + r * (SQRT5 + r * 5.0 / 3.0)) * e
            G = np.where(clamped, 0.0, kv * (5.0 / 3.0) * (1.0 + SQRT5 * r) * e)
        K = K0 + self.noise * np.eye(n)
        L = cholesky_nan(K)
        grad = np.full(self.num_hyperparams, np.nan)
        if not np.all(np.isfinite(L)):
            return float("nan"), grad
        alpha = cho_solve_lower(L, self.train_y)
        mll = float(-0.5 * (self.train_y.T @ alpha).item() - np.sum(np.log(np.diag(L))) - 0.5 * n * LOG_2PI)
        Linv = sla.solve_triangular(L, np.eye(n), lower=True, check_finite=False)
        W = alpha @ alpha.T - Linv.T @ Linv
        WG = W * G
        g = np.zeros(self.num_hyperparams)
        for j in range(d):
            diff = xs[:, j][:, None] - xs[:, j][None, :]
            g[j] = 0.5 * np.sum(WG * diff * diff)
        if not self.fixed_kernel_variance:
            g[d] = 0.5 * np.sum(W * K0)
        lp, gp_ = self.log_prior_and_grad(log_params)
        return -(mll + lp), -(g + gp_)

    # -- prediction ----------------------------------------------------------------------------
    def _k12(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        return self.kernel(self.train_x, x, self.lengthscales, self.kernel_variance, self.noise, False)

    def predict_mean_batched(self, x):
        """BOBE/gp.py:450-457,468-470 -- un-standardised mean."""
        k12 = self._k12(x)
        return (k12.T @ self.alphas).ravel() * self.y_std + self.y_mean

    def _raw_var(self, x):
        k12 = self._k12(x)
        vv = sla.solve_triangular(self.cholesky, k12, lower=True, check_finite=False)
        k22 = self.kernel_variance + self.noise  # kernel_diag(..., include_noise=True), gp.py:463
        return k22 - np.sum(vv * vv, axis=0)

    def predict_var_batched(self, x):
        """BOBE/gp.py:459-466,472-474 -- clip(var, 1e-12, None) (NaN propagates) times y_std^2."""
        var = self._raw_var(x)
        var = np.where(var < SAFE_NOISE_FLOOR, SAFE_NOISE_FLOOR, var)  # clip keeps NaN as NaN
        return self.y_std**2 * var

    def predict_batched(self, x):
        """BOBE/gp.py:476-493 -- standardised (mean, var), NaN and <1e-12 -> 1e-12; var has shape (M,1)."""
        k12 = self._k12(x)
        mean = (k12.T @ self.alphas).ravel()
        var = self._raw_var(x)
        var = np.where(np.isnan(var), SAFE_NOISE_FLOOR, var)
        var = np.where(var < SAFE_NOISE_FLOOR, SAFE_NOISE_FLOOR, var)
        return mean, var.reshape(-1, 1)

    def predict_mean_single(self, x):
        return float(self.predict_mean_batched(np.atleast_2d(x))[0])

    def predict_var_single(self, x):
        return float(self.predict_var_batched(np.atleast_2d(x))[0])

    def predict_single(self, x):
        m, v = self.predict_batched(np.atleast_2d(x))
        return float(m[0]), v[0]

    def predict_grad_batched(self, x, standardised=False):
        """(mean, var, dmean/dx, dvar/dx): the reverse-mode derivative of predict_mean_single / predict_var_single
        (standardised=False, BOBE/gp.py:450-466) or predict_single (standardised=True, BOBE/gp.py:476-489) with
        respect to the query point, as jax.grad yields at BOBE/samplers.py:268-285 and BOBE/optim.py:118,309.

          dk(x, x_j)/dx = -G_j (x - x_j) / l^2,  G = k (RBF) or kv 5/3 (1 + sqrt5 r) exp(-sqrt5 r) (Matern; 0 where the
          1e-30 clamp of gp.py:162 is active);  dmean = sum_j alpha_j dk_j;  dvar = -2 sum_j (K^-1 k*)_j dk_j, zero
          where the clip / where floor of gp.py:465,487-488 is active.  Pinned by central differences in
          tests/test_oracle.py.
        """
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        ls, kv = self.lengthscales, self.kernel_variance
        xs, qs = self.train_x / ls, x / ls
        q = dist_sq(qs, xs)  # (M, n)
        if self.kernel_name == "rbf":
            k = kv * np.exp(-0.5 * q)
            G = k
        else:
            clamped = q < 1e-30
            r = np.sqrt(np.where(clamped, 1e-30, q))
            e = np.exp(-SQRT5 * r)
            k = kv * (1.0 + r * (SQRT5 + r * 5.0 / 3.0)) * e
            G = np.where(clamped, 0.0, kv * (5.0 / 3.0) * (1.0 + SQRT5 * r) * e)
        alpha = self.alphas.ravel()
        mean = k @ alpha
        vv = sla.solve_triangular(self.cholesky, k.T, lower=True, check_finite=False)
        var = (kv + self.noise) - np.sum(vv * vv, axis=0)
        w = sla.solve_triangular(self.cholesky.T, vv, lower=False, check_finite=False).T  # (M, n) = K^-1 k*
        diff = (x[:, None, :] - self.train_x[None, :, :]) / (ls * ls)  # (M, n, d)
        dmean = -np.einsum("mj,mjk->mk", alpha[None, :] * G, diff)
        dvar = 2.0 * np.einsum("mj,mjk->mk", w * G, diff)
        if standardised:
            var = np.where(np.isnan(var), SAFE_NOISE_FLOOR, var)
        floored = ~(var > SAFE_NOISE_FLOOR)
        var = np.where(var < SAFE_NOISE_FLOOR, SAFE_NOISE_FLOOR, var)
        dvar = np.where(floored[:, None], 0.0, dvar)
        if standardised:
            return mean, var, dmean, dvar
        return mean * self.y_std + self.y_mean, var * self.y_std**2, dmean * self.y_std, dvar * self.y_std**2

    # -- update --------------------------------------------------------------------------------
    def update(self, new_x, new_y):
        """BOBE/gp.py:495-541 -- dedupe (isclose atol 1e-6 rtol 1e-4 on all dims), re-standardise, re-factor."""
        new_x = np.atleast_2d(np.asarray(new_x, dtype=np.float64))
        new_y = np.atleast_2d(np.asarray(new_y, dtype=np.float64))
        pts, vals = [], []
        for i in range(new_x.shape[0]):
            if np.any(np.all(np.isclose(self.train_x, new_x[i], atol=1e-6, rtol=1e-4), axis=1)):
                continue
            pts.append(new_x[i])
            vals.append(new_y[i])
        if pts:
            self.train_x = np.vstack([self.train_x, np.array(pts)])
            y_orig = np.vstack([self.train_y * self.y_std + self.y_mean, np.array(vals).reshape(len(vals), -1)])
            self.y_mean = float(np.mean(y_orig))
            self.y_std = float(np.std(y_orig))
            if self.y_std == 0:
                self.y_std = 1.0
            self.train_y = (y_orig - self.y_mean) / self.y_std
            self.recompute_cholesky()

    # -- fantasy variance ----------------------------------------------------------------------
    def fantasy_var(self, new_x, mc_points, k_train_mc):
        """BOBE/gp.py:552-576, literally: rank-1 append then a full (n+1)x(n+1) TRSM against all MC columns."""
        new_x = np.atleast_2d(np.asarray(new_x, dtype=np.float64))
        k = self.kernel(self.train_x, new_x, self.lengthscales, self.kernel_variance, self.noise, False).ravel()
        k_self = self.kernel_variance + self.noise
        k11_cho = fast_update_cholesky(self.cholesky, k, k_self)
        k_new_mc = self.kernel(new_x, mc_points, self.lengthscales, self.kernel_variance, self.noise, False)
        k12 = np.vstack([k_train_mc, k_new_mc])
        k22 = kernel_diag(mc_points, self.kernel_variance, self.noise, True)
        if np.isfinite(k11_cho[-1, -1]) and k11_cho[-1, -1] != 0.0:
            vv = sla.solve_triangular(k11_cho, k12, lower=True, check_finite=False)
        else:  # NaN / zero pivot: LAPACK would raise or warn; XLA's trsm just propagates NaN/inf
            vv = np.empty_like(k12)
            vv[:-1] = sla.solve_triangular(self.cholesky, k_train_mc, lower=True, check_finite=False)
            with np.errstate(all="ignore"):
                vv[-1] = (k_new_mc.ravel() - k11_cho[-1, :-1] @ vv[:-1]) / k11_cho[-1, -1]
        with np.errstate(all="ignore"):
            var = k22 - np.sum(vv * vv, axis=0)
        var = np.where(np.isnan(var), SAFE_NOISE_FLOOR, var)
        var = np.where(var < SAFE_NOISE_FLOOR, SAFE_NOISE_FLOOR, var)
        return var * self.y_std**2

    def fantasy_var_shared(self, cand_x, mc_points):
        """Same quantity for many candidates via the shared V = L^-1 K(X,MC).

        Algebraically identical to ``fantasy_var`` column by column; used as the CPU baseline for the
        WIPV workload and to check the algebra the CUDA kernel uses.  Returns (C, n_mc).
        """
        cand_x = np.atleast_2d(np.asarray(cand_x, dtype=np.float64))
        kk = self.kernel_variance + self.noise
        V = sla.solve_triangular(self.cholesky, self._k12(mc_points), lower=True, check_finite=False)
        Vc = sla.solve_triangular(self.cholesky, self._k12(cand_x), lower=True, check_finite=False)
        base = kk - np.sum(V * V, axis=0)
        with np.errstate(all="ignore"):
            delta2 = kk - np.sum(Vc * Vc, axis=0)
            delta = np.sqrt(delta2)
            kc = self.kernel(cand_x, mc_points, self.lengthscales, self.kernel_variance, self.noise, False)
            w = (kc - Vc.T @ V) / delta[:, None]
            var = base[None, :] - w * w
        var = np.where(np.isnan(var), SAFE_NOISE_FLOOR, var)
        var = np.where(var < SAFE_NOISE_FLOOR, SAFE_NOISE_FLOOR, var)
        return var * self.y_std**2


# ----------------------------------------------------------------------------------------------
# acquisition arithmetic  (BOBE/acquisition.py:21-75,226-253,318-330,438-465)
# ----------------------------------------------------------------------------------------------
def _log_phi(u):
    """BOBE/acquisition.py:25-27."""
    return -0.5 * (u**2 + LOG_2PI)


def _ei_helper(u):
    """BOBE/acquisition.py:29-31 -- phi(u) + u Phi(u)."""
    return np.exp(-0.5 * u * u) / math.sqrt(2 * math.pi) + u * ssp.ndtr(u)


def _log1mexp(x):
    """tfp.math.log1mexp: log(1 - exp(-|x|)), switch at log 2."""
    x = np.abs(x)
    with np.errstate(all="ignore"):
        return np.where(x < math.log(2.0), np.log(-np.expm1(-x)), np.log1p(-np.exp(-x)))


def _log_abs_u_Phi_div_phi(u):
    """BOBE/acquisition.py:33-42."""
    return np.log(np.abs(u) * ssp.erfcx(-u / SQRT2)) + 0.5 * math.log(math.pi / 2.0)


def log_ei_helper(u):
    """BOBE/acquisition.py:44-75 (float64 branch constants)."""
    u = np.asarray(u, dtype=np.float64)
    bound, neg_inv_sqrt_eps = -1.0, -1e6
    u_upper = np.where(u < bound, bound, u)
    with np.errstate(all="ignore"):
        log_ei_upper = np.log(_ei_helper(u_upper))
        u_lower = np.where(u > bound, bound, u)
        u_eps = np.where(u_lower < neg_inv_sqrt_eps, neg_inv_sqrt_eps, u_lower)
        w = _log_abs_u_Phi_div_phi(u_eps)
        second = np.where(u > neg_inv_sqrt_eps, _log1mexp(w), -2.0 * np.log(np.abs(u_lower)))
        log_ei_lower = _log_phi(u) + second
    return np.where(u > bound, log_ei_upper, log_ei_lower)


def ei_values(mu, var, best_y, zeta):
    """BOBE/acquisition.py:226-253 -- returns the NEGATED EI the optimiser minimises; mu/var standardised."""
    var = np.maximum(np.asarray(var, dtype=np.float64).ravel(), 1e-20)
    sigma = np.sqrt(var)
    u = (np.asarray(mu) - zeta - best_y) / sigma
    return -(_ei_helper(u) * sigma)


def logei_values(mu, var, best_y, zeta):
    """BOBE/acquisition.py:318-330 -- negated log-EI."""
    var = np.maximum(np.asarray(var, dtype=np.float64).ravel(), 1e-18)
    sigma = np.sqrt(var)
    u = (np.asarray(mu) - zeta - best_y) / sigma
    return -(log_ei_helper(u) + np.log(sigma))


def svm_decision(x, support_vectors, dual_coef, intercept, gamma):
    """BOBE/clf.py:188-209 -- RBF-SVM decision function sum_j dual_j exp(-gamma |sv_j - x|^2) + intercept, (M,)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    return np.exp(-gamma * dist_sq(x, np.asarray(support_vectors))) @ np.asarray(dual_coef).ravel() + intercept


def clf_masked_predict(gp: "OracleGP", x, clf_params, minus_inf=-1e5, standardised=False):
    """BOBE/clf_gp.py:173-205 -- where(clf_probs >= threshold, value, fill) with probs = (decision >= 0)."""
    dec = svm_decision(x, clf_params['support_vectors'], clf_params['dual_coef'], clf_params['intercept'],
                       clf_params['gamma_eff'])
    ok = dec >= 0
    if standardised:
        m, v = gp.predict_batched(x)
        v = v.ravel()
    else:
        m, v = gp.predict_mean_batched(x), gp.predict_var_batched(x)
    return np.where(ok, m, minus_inf), np.where(ok, v, SAFE_NOISE_FLOOR), dec


def wipv_values(gp: OracleGP, cand_x, mc_points, std=False):
    """BOBE/acquisition.py:438-440,463-465 for a set of candidates: mean_j fantasy_var (or of its sqrt)."""
    var = gp.fantasy_var_shared(cand_x, mc_points)
    return np.mean(np.sqrt(var) if std else var, axis=1)


def _kernel_and_gradcoef(gp: OracleGP, xa, xb):
    """k(xa, xb) and G with dk(x, p)/dx = -G (x - p) / l^2 (G = k for RBF; Matern-5/2: kv 5/3 (1 + sqrt5 r) exp(-sqrt5 r),
    0 where the 1e-30 clamp of BOBE/gp.py:162 is active)."""
    ls, kv = gp.lengthscales, gp.kernel_variance
    q = dist_sq(xa / ls, xb / ls)
    if gp.kernel_name == "rbf":
        k = kv * np.exp(-0.5 * q)
        return k, k
    clamped = q < 1e-30
    r = np.sqrt(np.where(clamped, 1e-30, q))
    e = np.exp(-SQRT5 * r)
    return kv * (1.0 + r * (SQRT5 + r * 5.0 / 3.0)) * e, np.where(clamped, 0.0, kv * (5.0 / 3.0) * (1.0 + SQRT5 * r) * e)


def wipv_values_and_grad(gp: OracleGP, cand_x, mc_points, std=False):
    """WIPV / WIPStd (BOBE/acquisition.py:438-440,463-465) and their gradient with respect to the candidate point:
    what ``jax.value_and_grad(self.fun)`` yields in the n <= 500 polish of BOBE/acquisition.py:400-412 (through
    BOBE/optim.py:118,309), i.e. the reverse-mode derivative of GP.fantasy_var (BOBE/gp.py:552-576) in ``new_x``.

    With v = L^-1 k(X,x), delta2 = k** - v.v, V_j = L^-1 k(X,mc_j), t_j = k(x,mc_j) - v.V_j (the posterior covariance
    of x and mc_j) and s_j = k** - |V_j|^2 - t_j^2 / delta2:
        ds_j/dx = -2 t_j t_j'/delta2 + t_j^2 delta2'/delta2^2,
        t_j'    = dk(x,mc_j)/dx - sum_i dk(x,X_i)/dx (K^-1 k(X,mc_j))_i,     delta2' = -2 sum_i (K^-1 k(X,x))_i dk(x,X_i)/dx,
    zero where the NaN / 1e-12 floor of gp.py:574-575 is active.  Pinned by central differences in tests/test_oracle.py.
    Returns (values (C,), gradients (C, d))."""
    cand_x = np.atleast_2d(np.asarray(cand_x, dtype=np.float64))
    mc_points = np.atleast_2d(np.asarray(mc_points, dtype=np.float64))
    L, kk, scale = gp.cholesky, gp.kernel_variance + gp.noise, gp.y_std**2
    ls2 = gp.lengthscales**2
    V = sla.solve_triangular(L, gp._k12(mc_points), lower=True, check_finite=False)  # (n, n_mc)
    base = kk - np.sum(V * V, axis=0)
    Wmc = sla.solve_triangular(L.T, V, lower=False, check_finite=False)  # K^-1 k(X, MC)
    vals, grads = [], []
    for x in cand_x:
        x = x[None, :]
        kx, Gx = _kernel_and_gradcoef(gp, x, gp.train_x)  # (1, n)
        kc, Gc = _kernel_and_gradcoef(gp, x, mc_points)  # (1, n_mc)
        v = sla.solve_triangular(L, kx.T, lower=True, check_finite=False)  # (n, 1)
        w = sla.solve_triangular(L.T, v, lower=False, check_finite=False).ravel()
        with np.errstate(all="ignore"):
            delta2 = kk - float(np.sum(v * v))
            if delta2 < 0:
                delta2 = np.nan  # sqrt of a negative pivot, gp.py:187
            t = (kc - v.T @ V).ravel()
            s = base - t * t / delta2
        floored = ~(s >= SAFE_NOISE_FLOOR)  # NaN or below the floor
        s = np.where(floored, SAFE_NOISE_FLOOR, s)
        val = scale * s
        c = np.full_like(s, scale) if not std else scale / (2.0 * np.sqrt(val))
        c = np.where(floored, 0.0, c)
        with np.errstate(all="ignore"):
            a = np.where(floored, 0.0, -2.0 * c * t / delta2)
            b = float(np.sum(np.where(floored, 0.0, c * t * t / delta2**2)))
        dk_mc = -(Gc.ravel()[:, None]) * (x - mc_points) / ls2  # (n_mc, d) = dk(x, mc_j)/dx
        dk_tr = -(Gx.ravel()[:, None]) * (x - gp.train_x) / ls2  # (n, d)   = dk(x, X_i)/dx
        e = -(Wmc @ a) - 2.0 * b * w  # coefficient of dk(x, X_i)/dx
        grads.append((a @ dk_mc + e @ dk_tr) / mc_points.shape[0])
        vals.append(np.mean(np.sqrt(val) if std else val))
    return np.array(vals), np.array(grads)


# ----------------------------------------------------------------------------------------------
# synthetic workloads -- shared by tests and bench so that both sides see the same inputs
# ----------------------------------------------------------------------------------------------
def synthetic_training_set(n, d, seed=0):
    """X ~ U[0,1]^(n x d), y = -1/2 sum((x-0.5)/0.15)^2 (Gaussian log-likelihood shape)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 1.0, (n, d))
    y = -0.5 * np.sum(((X - 0.5) / 0.15) ** 2, axis=1, keepdims=True)
    return X, y


def synthetic_queries(m, d, seed=1):
    return np.random.default_rng(seed).uniform(0.0, 1.0, (m, d))


def synthetic_restarts(gp, n_restarts, seed=2):
    """BOBE/pool.py:277-286 -- row 0 = log(current hp), the rest uniform in the log-bounds."""
    rng = np.random.default_rng(seed)
    init = np.log(gp.get_hyperparams())
    if n_restarts > 1:
        x0r = rng.uniform(gp.hyperparam_bounds[0], gp.hyperparam_bounds[1], size=(n_restarts - 1, gp.num_hyperparams))
        return np.vstack([init, x0r])
    return np.atleast_2d(init)
