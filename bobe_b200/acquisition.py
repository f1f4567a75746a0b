"""Drop-in for the acquisition interface of ``BOBE/acquisition.py`` on the B200 GP hot path.

Same class names, ``fun`` / ``get_next_point`` / ``get_next_batch`` signatures and return conventions as the
reference (BOBE/acquisition.py:79-489).  Differences, all forced by the absence of JAX autodiff here:
  * gradients for the polish step are ANALYTIC (``bobe_predict_grad`` for EI / LogEI, ``bobe_fantasy_var_grad`` for
    WIPV / WIPStd), evaluated in ONE batched device call per optimiser step for all restarts, instead of
    ``jax.value_and_grad``; batched central differences remain the default for user-defined acquisitions;
  * the candidate sweep of WIPV/WIPStd (``lax.map`` over MC points, :390-394) is one fused device call that
    shares V = L^-1 K(X, MC) between all candidates.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
from scipy.stats import qmc

from . import ops
from .gp import GP, _to_dev
from .optim import optimize_optax, optimize_optax_vmap, optimize_scipy

log = logging.getLogger("bobe_b200.acq")

FD_STEP = 1e-6


def _phi(u):
    return np.exp(-0.5 * u * u) * 0.3989422804014327


def _ndtr(u):
    from scipy.special import ndtr
    return ndtr(u)


def _ei_ratios(u):
    """(phi/h, Phi/h) with h(u) = phi(u) + u Phi(u), stable in both tails.

    u > -1: direct.  u <= -1: through R = Phi/phi = sqrt(pi/2) erfcx(-u/sqrt2), h/phi = 1 + u R; below u = -50 the
    cancellation in 1 + u R (-> 1/u^2) is replaced by its asymptotic series 1/u^2 - 3/u^4 + 15/u^6 - 105/u^8."""
    from scipy.special import erfcx, ndtr
    u = np.asarray(u, dtype=np.float64)
    hi = u > -1.0
    uh = np.where(hi, u, 0.0)
    h = _phi(uh) + uh * ndtr(uh)
    a_hi, b_hi = _phi(uh) / h, ndtr(uh) / h
    ul = np.where(hi, -1.0, u)
    R = 1.2533141373155003 * erfcx(-ul * 0.7071067811865476)
    iu2 = 1.0 / (ul * ul)
    den = np.where(ul < -50.0, iu2 * (1.0 + iu2 * (-3.0 + iu2 * (15.0 - 105.0 * iu2))), 1.0 + ul * R)
    return np.where(hi, a_hi, 1.0 / den), np.where(hi, b_hi, R / den)


def _fd_batched(fun_batched, lo=0.0, hi=1.0, h=FD_STEP):
    """(R, d) -> values (R,), gradients (R, d) by central differences, one batched call."""

    def vg(xs):
        xs = np.atleast_2d(np.asarray(xs, dtype=np.float64))
        R, d = xs.shape
        pts = np.repeat(xs[:, None, :], 2 * d + 1, axis=1)  # (R, 2d+1, d)
        hp = np.minimum(h, hi - xs)  # one-sided at the box faces
        hm = np.minimum(h, xs - lo)
        for j in range(d):
            pts[:, 1 + 2 * j, j] += hp[:, j]
            pts[:, 2 + 2 * j, j] -= hm[:, j]
        f = np.asarray(fun_batched(pts.reshape(-1, d)), dtype=np.float64).reshape(R, 2 * d + 1)
        g = np.zeros((R, d))
        for j in range(d):
            den = hp[:, j] + hm[:, j]
            g[:, j] = np.where(den > 0, (f[:, 1 + 2 * j] - f[:, 2 + 2 * j]) / np.where(den > 0, den, 1.0), 0.0)
        return f[:, 0], g

    return vg


class AcquisitionFunction:
    """BOBE/acquisition.py:79-196."""

    name: str = "BaseAcquisitionFunction"

    def __init__(self, optimizer: str = "scipy", optimizer_options: Optional[Dict[str, Any]] = {}):
        self.optimizer = optimizer
        self.optimizer_options = optimizer_options
        self.acq_optimize = optimize_scipy if self.optimizer == "scipy" else optimize_optax

    def fun(self, x):
        raise NotImplementedError

    def get_next_point(self, gp: GP, acq_kwargs: Dict[str, Any] = {}, maxiter: int = 500, n_restarts: int = 8,
                       verbose: bool = True, early_stop_patience: int = 25, rng=None) -> Tuple[np.ndarray, float]:
        raise NotImplementedError("Base class get_next() not implemented")

    def _optimize(self, fun_batched, x0, gp, maxiter, n_restarts, verbose, vg_batched=None):
        """``vg_batched``: analytic (values, gradients) for (R, d) points; default = batched central differences."""
        x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
        vg_b = vg_batched if vg_batched is not None else _fd_batched(fun_batched)
        opts = dict(self.optimizer_options)
        return self.acq_optimize(fun=None, num_params=gp.ndim, x0=x0, bounds=[0, 1], optimizer_options=opts,
                                 maxiter=maxiter, n_restarts=min(n_restarts, x0.shape[0]), verbose=verbose,
                                 batched_value_and_grad=vg_b)

    def get_next_batch(self, gp: GP, n_batch: int = 1, acq_kwargs: Dict[str, Any] = {}, maxiter: int = 500,
                       n_restarts: int = 8, verbose: bool = True, early_stop_patience: int = 25,
                       rng=None) -> Tuple[np.ndarray, np.ndarray]:
        """BOBE/acquisition.py:147-196 -- greedy batch with a kriging-believer dummy GP."""
        rng = rng if rng is not None else np.random.default_rng()
        x_batch, acq_vals = [], []
        x_next, acq_val_next = self.get_next_point(gp, acq_kwargs=acq_kwargs, maxiter=maxiter, n_restarts=n_restarts,
                                                   verbose=verbose, early_stop_patience=early_stop_patience, rng=rng)
        x_batch.append(x_next)
        acq_vals.append(acq_val_next)
        if n_batch > 1:
            dummy_gp = GP(train_x=gp.train_x, train_y=gp.train_y * gp.y_std + gp.y_mean, noise=gp.noise,
                          kernel=gp.kernel_name, lengthscales=gp.lengthscales, kernel_variance=gp.kernel_variance)
            dummy_gp.update(x_next, dummy_gp.predict_mean_single(x_next))
            for _ in range(1, n_batch):
                x_next, acq_val_next = self.get_next_point(dummy_gp, acq_kwargs=acq_kwargs, maxiter=maxiter,
                                                           n_restarts=n_restarts, verbose=verbose,
                                                           early_stop_patience=early_stop_patience, rng=rng)
                x_batch.append(x_next)
                acq_vals.append(acq_val_next)
                dummy_gp.update(x_next, dummy_gp.predict_mean_single(x_next))
        return np.array(x_batch), np.array(acq_vals)


class EI(AcquisitionFunction):
    """BOBE/acquisition.py:199-291."""

    name: str = "EI"
    _which = "ei"

    def __init__(self, optimizer: str = "scipy", optimizer_options: Optional[Dict[str, Any]] = {}):
        super().__init__(optimizer=optimizer, optimizer_options=optimizer_options)
        if optimizer == 'optax':
            self.acq_optimize = optimize_optax_vmap

    def fun_batched(self, x, gp, best_y, zeta):
        """Negated (log-)EI at (M, d) points in one device pass: predict_batched + the EI epilogue kernel."""
        as_t = isinstance(x, torch.Tensor)
        gp._ensure_factor()
        xq = _to_dev(x, gp.device)
        if xq.dim() == 1:
            xq = xq[None, :]
        if xq.shape[1] != gp.ndim:
            raise ValueError(f"query points must have {gp.ndim} columns")
        # gp.predict_single semantics (BOBE/acquisition.py:246,323), through the GP's own device path so that a
        # GPwithClassifier applies its feasibility mask (BOBE/clf_gp.py:197-205)
        mean, var = gp._predict_dev(xq, True, True, True)
        out = ops.acq_ei(self._which, mean, var, float(best_y), float(zeta))
        return out if as_t else out.cpu().numpy()

    def fun(self, x, gp, best_y, zeta):
        """BOBE/acquisition.py:226-253 -- negative EI (the optimiser minimises)."""
        return self.fun_batched(np.atleast_2d(np.asarray(x, dtype=np.float64)), gp, best_y, zeta)[0]

    def value_and_grad_batched(self, x, gp, best_y, zeta):
        """Negated (log-)EI and its gradient w.r.t. x at (R, d) points -- ``jax.value_and_grad(self.fun)`` of
        BOBE/optim.py:118,309, analytically: one ``bobe_predict_grad`` pass (standardised mean / variance and their
        input gradients) + the closed-form partials of the epilogue (chain rule on the host, O(R d))."""
        xs = np.atleast_2d(np.asarray(x, dtype=np.float64))
        mu, var, dmu, dvar = gp.predict_grad_batched(xs, standardised=True)
        floor = 1e-20 if self._which == "ei" else 1e-18  # jnp.clip(var, a_min=...), acquisition.py:247,324
        clipped = var < floor
        v = np.maximum(var, floor)
        sigma = np.sqrt(v)
        u = (mu - zeta - best_y) / sigma
        val = ops.acq_ei(self._which, torch.as_tensor(mu, device=gp.device), torch.as_tensor(var, device=gp.device),
                         float(best_y), float(zeta)).cpu().numpy()
        if self._which == "ei":  # d(-EI)/dmu = -Phi(u), d(-EI)/dsigma = -phi(u)
            d_mu = -_ndtr(u)
            d_v = -_phi(u) / (2.0 * sigma)
        else:  # log EI = log sigma + log h(u), h = phi + u Phi:  d/dmu = (Phi/h)/sigma,  d/dv = (phi/h)/(2 v)
            phi_over_h, Phi_over_h = _ei_ratios(u)
            d_mu = -Phi_over_h / sigma
            d_v = -phi_over_h / (2.0 * v)
        d_v = np.where(clipped, 0.0, d_v)
        # u depends on v through sigma as well; the partials above already are the total derivatives in (mu, v)
        grad = d_mu[:, None] * dmu + d_v[:, None] * dvar
        return val, grad

    def get_next_point(self, gp, acq_kwargs, maxiter: int = 250, n_restarts: int = 20, verbose: bool = True,
                       early_stop_patience: int = 25, rng=None):
        """BOBE/acquisition.py:255-291."""
        rng = rng if rng is not None else np.random.default_rng()
        zeta = acq_kwargs.get('zeta', 0.)
        best_y = acq_kwargs.get('best_y', max(gp.train_y.flatten()))
        best_x = gp.train_x[np.argmax(gp.train_y)]
        if n_restarts > 1:
            n_random_restarts = int(n_restarts / 2)
            x0_acq = np.vstack([gp.get_random_point(rng, nstd=5) for _ in range(n_random_restarts)])
            n_best_restarts = n_restarts - n_random_restarts
            x0_acq = np.vstack([x0_acq, np.full((n_best_restarts, gp.ndim), best_x)])
        else:
            x0_acq = np.atleast_2d(best_x)
        jitter = rng.normal(0., 0.005, size=x0_acq.shape)
        x0_acq = np.clip(x0_acq + jitter, 0., 1.)
        pts, vals = self._optimize(lambda xs: self.fun_batched(xs, gp, best_y, zeta), x0_acq, gp, maxiter, n_restarts,
                                   verbose, vg_batched=lambda xs: self.value_and_grad_batched(xs, gp, best_y, zeta))
        return pts, -vals  # we minimise -EI


class LogEI(EI):
    """BOBE/acquisition.py:293-330."""

    name: str = "LogEI"
    _which = "logei"


class WeightedIntegratedPosteriorBase(AcquisitionFunction):
    """BOBE/acquisition.py:333-412."""

    _std = False

    def fun_batched(self, x, gp, mc_points=None, k_train_mc=None):
        return gp.fantasy_acquisition(mc_points, np.atleast_2d(x) if not isinstance(x, torch.Tensor) else x, self._std)

    def fun(self, x, gp, mc_points=None, k_train_mc=None):
        return self.fun_batched(np.atleast_2d(np.asarray(x, dtype=np.float64)), gp, mc_points=mc_points)[0]

    def value_and_grad_batched(self, x, gp, mc_points=None, k_train_mc=None):
        """``jax.value_and_grad(self.fun)`` of BOBE/optim.py:118,309 at (R, d) points, analytically on the device."""
        return gp.fantasy_acquisition_value_and_grad(mc_points, np.atleast_2d(np.asarray(x, dtype=np.float64)), self._std)

    def get_next_point(self, gp, acq_kwargs, maxiter: int = 100, n_restarts: int = 1, verbose: bool = True,
                       early_stop_patience: int = 25, rng=None):
        mc_samples = acq_kwargs.get('mc_samples')
        mc_points_size = acq_kwargs.get('mc_points_size', 128)
        mc_points = np.asarray(get_mc_points(mc_samples, mc_points_size=mc_points_size, rng=rng), dtype=np.float64)
        # every MC point is a candidate (BOBE/acquisition.py:390-397): one fused call
        acq_vals = gp.fantasy_acquisition(mc_points, None, self._std)
        i = int(np.argmin(acq_vals))
        acq_val_min = float(acq_vals[i])
        x0_acq = mc_points[i]
        if gp.train_x.shape[0] > 500:  # BOBE/acquisition.py:400-401
            return x0_acq, acq_val_min
        return self._optimize(lambda xs: self.fun_batched(xs, gp, mc_points=mc_points), x0_acq, gp, maxiter,
                              n_restarts, verbose,
                              vg_batched=lambda xs: self.value_and_grad_batched(xs, gp, mc_points=mc_points))


class WIPV(WeightedIntegratedPosteriorBase):
    """BOBE/acquisition.py:415-440."""

    name: str = "WIPV"
    _std = False


class WIPStd(WeightedIntegratedPosteriorBase):
    """BOBE/acquisition.py:443-465."""

    name: str = "WIPStd"
    _std = True


def get_mc_samples(gp: GP, warmup_steps=512, num_samples=1024, thinning=4, method="NUTS", num_chains=4, np_rng=None,
                   rng_key=None):
    """BOBE/acquisition.py:468-482.  Only the sampler-free 'uniform' method is built in; NUTS / NS live in the
    reference's samplers.py (numpyro / dynesty), which is outside the hot-path scope."""
    if method == 'uniform':
        points = qmc.Sobol(gp.ndim, scramble=True, rng=np_rng).random(num_samples)
        return {'x': points}
    if method in ('NUTS', 'NS'):
        raise NotImplementedError(f"method={method!r} needs the reference's samplers (numpyro / dynesty); "
                                  "pass mc_samples={'x': ...} from your sampler, or use method='uniform'")
    raise ValueError(f"Unknown method {method} for sampling GP")


def get_mc_points(mc_samples, mc_points_size=128, rng=None):
    """BOBE/acquisition.py:485-489."""
    mc_size = max(mc_samples['x'].shape[0], mc_points_size)
    rng = rng if rng is not None else np.random.default_rng()
    idxs = rng.choice(mc_size, size=mc_points_size, replace=False)
    return mc_samples['x'][idxs]


ACQUISITIONS = {"wipv": WIPV, "ei": EI, "logei": LogEI, "wipstd": WIPStd}  # registry at BOBE/bo.py:22
