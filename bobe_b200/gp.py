"""Drop-in for ``BOBE/gp.py``: the ``GP`` class and module-level kernel functions on the B200 hot path.

Same names, argument meaning and error behaviour as the reference (BOBE/gp.py:80-772); the arithmetic is
done by hand-written sm_100a kernels through the C-ABI in ``include/bobe_b200.h``.  Host arrays in and out are
NumPy (the reference uses ``jnp`` arrays); CUDA ``torch`` tensors are accepted anywhere an array is and are
then returned as CUDA tensors without a host round trip.  There is no CPU path: a GP can be constructed and
its host-side bookkeeping exercised without a GPU, but any arithmetic raises unless CUDA is available.
"""
from __future__ import annotations

import functools
import logging
from typing import List

import numpy as np
import torch

from . import ops
from . import priors as _pr
from .optim import optimize_optax, optimize_scipy

log = logging.getLogger("bobe_b200.gp")

safe_noise_floor = 1e-12  # BOBE/gp.py:16


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("bobe_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _is_t(x) -> bool:
    return isinstance(x, torch.Tensor)


def _to_dev(x, device) -> torch.Tensor:
    if _is_t(x):  # pinned CPU tensors copy asynchronously on the current stream
        return x.to(device=device, dtype=torch.float64, non_blocking=True)
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64)), device=device)


# ---- module-level kernel functions (BOBE/gp.py:80-168) -----------------------------------------------------
def _kernel_call(kind, xa, xb, lengthscales, kernel_variance, noise, include_noise=True, device=None):
    as_t = _is_t(xa) or _is_t(xb)
    dev = xa.device if _is_t(xa) and xa.is_cuda else (xb.device if _is_t(xb) and xb.is_cuda else _dev(device))
    xa_d, xb_d = _to_dev(xa, dev), _to_dev(xb, dev)
    if xa_d.dim() != 2 or xb_d.dim() != 2:
        raise ValueError("kernel inputs must be 2D (n, d)")
    ls_d = _to_dev(lengthscales, dev).reshape(-1)
    if ls_d.numel() == 1 and xa_d.shape[1] > 1:
        ls_d = ls_d.expand(xa_d.shape[1]).contiguous()
    k = ops.kernel_matrix(kind, xa_d, xb_d, ls_d, float(kernel_variance), float(noise), bool(include_noise))
    return k if as_t else k.cpu().numpy()


def rbf_kernel(xa, xb, lengthscales, kernel_variance, noise, include_noise=True):
    """BOBE/gp.py:124-154."""
    return _kernel_call("rbf", xa, xb, lengthscales, kernel_variance, noise, include_noise)


def matern_kernel(xa, xb, lengthscales, kernel_variance, noise, include_noise=True):
    """BOBE/gp.py:156-168."""
    return _kernel_call("matern", xa, xb, lengthscales, kernel_variance, noise, include_noise)


def kernel_diag(x, kernel_variance, noise, include_noise=True):
    """BOBE/gp.py:98-122 -- a constant vector; no device work."""
    n = x.shape[0]
    diag = kernel_variance * np.ones(n)
    if include_noise:
        diag = diag + noise
    return diag


def dist_sq(x, y):
    """BOBE/gp.py:80-96 -- squared distances by direct differences (``bobe_dist_sq``)."""
    as_t = _is_t(x)
    dev = x.device if as_t and x.is_cuda else _dev()
    out = ops.dist_sq(_to_dev(np.atleast_2d(x) if not as_t else x, dev), _to_dev(np.atleast_2d(y) if not _is_t(y) else y, dev))
    return out if as_t else out.cpu().numpy()


def gp_mll(k, train_y, num_points):
    """BOBE/gp.py:170-178 -- log marginal likelihood of a caller-supplied kernel matrix (``bobe_cholesky_batched``):
    ``-1/2 y^T K^-1 y - sum log L_ii - n/2 log 2 pi``; NaN when K is not positive definite (jnp.linalg.cholesky)."""
    as_t = _is_t(k)
    dev = k.device if as_t and k.is_cuda else _dev()
    _, _, logdet, quad, _ = ops.cholesky_solve(_to_dev(k, dev), _to_dev(train_y, dev).reshape(-1))
    val = -0.5 * quad[0] - logdet[0] - 0.5 * float(num_points) * float(np.log(2.0 * np.pi))
    return val if as_t else float(val.item())


def fast_update_cholesky(L, k, k_self):
    """BOBE/gp.py:181-197."""
    as_t = _is_t(L)
    dev = L.device if as_t else _dev()
    out = ops.chol_append(_to_dev(L, dev), _to_dev(k, dev), float(k_self))
    return out if as_t else out.cpu().numpy()


class GP:
    """Gaussian-process surrogate with the reference's public surface (BOBE/gp.py:199-772)."""

    def __init__(self, train_x, train_y, noise=1e-8, kernel="rbf", optimizer="scipy", optimizer_options={},
                 kernel_variance_bounds=[1e-4, 1e8], lengthscale_bounds=[0.01, 5], lengthscales=None,
                 kernel_variance=None, kernel_variance_prior=None, lengthscale_prior=None, tausq=None,
                 tausq_bounds=[1e-4, 1e4], param_names: List[str] = None, device=None):
        # The device is fixed at construction: torch's current device is THREAD-local (new threads start on device 0),
        # and the lock-step optimisers / SurrogatePool call back into this object from worker threads.
        self._device_arg = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None)
        self._setup_training_data(train_x, train_y)
        self.param_names = param_names if param_names is not None else ['x_' + str(i) for i in range(self.ndim)]

        self.kernel_name = kernel if kernel == "rbf" else "matern"  # BOBE/gp.py:251-252
        # same call signature as the module-level kernels, but bound to this GP's device (worker threads of the
        # lock-step optimisers would otherwise land on torch's thread-local default device)
        self.kernel = functools.partial(_kernel_call, self.kernel_name, device=self._device_arg)
        self.lengthscales = (np.asarray(lengthscales, dtype=np.float64).reshape(-1) if lengthscales is not None
                             else np.ones(self.ndim))
        self.kernel_variance = float(kernel_variance) if kernel_variance is not None else 1.0
        self.noise = noise

        self.optimizer_method = optimizer  # BOBE/gp.py:263-268
        self.mll_optimize = optimize_scipy if optimizer == "scipy" else optimize_optax
        self.optimizer_options = optimizer_options

        self.lengthscale_bounds = lengthscale_bounds
        self.kernel_variance_bounds = kernel_variance_bounds
        self.tausq = tausq if tausq is not None else 1.0
        self.tausq_bounds = tausq_bounds

        self._setup_kernel_variance_prior(kernel_variance_prior)
        self._setup_lengthscale_prior(lengthscale_prior)
        self._setup_optimization_parameters()
        self._invalidate()
        if torch.cuda.is_available():  # the reference factorises in __init__ (BOBE/gp.py:258-260)
            self._ensure_factor()

    # ---- set-up (host) ---------------------------------------------------------------------------------
    def _setup_training_data(self, train_x, train_y):
        """BOBE/gp.py:283-307."""
        train_x = train_x.detach().cpu().numpy() if _is_t(train_x) else np.asarray(train_x, dtype=np.float64)
        train_y = train_y.detach().cpu().numpy() if _is_t(train_y) else np.asarray(train_y, dtype=np.float64)
        if train_x.shape[0] != train_y.shape[0]:
            raise ValueError("train_x and train_y must have the same number of points")
        if train_y.ndim != 2:
            train_y = train_y.reshape(-1, 1)
        if train_x.ndim != 2:
            raise ValueError("train_x must be 2D")
        self.ndim = train_x.shape[1]
        self.y_mean = float(np.mean(train_y)) if train_y.size > 0 else 0
        self.y_std = float(np.std(train_y)) if train_y.size > 0 else 1.0
        if self.y_std == 0:
            log.warning("Training targets have zero variance. Setting std to 1.0 to avoid division by zero.")
            self.y_std = 1.0
        self.train_x = np.array(train_x, dtype=np.float64)
        self.train_y = (train_y - self.y_mean) / self.y_std

    def _setup_kernel_variance_prior(self, kernel_variance_prior):
        """BOBE/gp.py:309-320."""
        self.kernel_variance_prior_spec = kernel_variance_prior
        if self.kernel_variance_prior_spec is None:
            self.kernel_variance_prior_spec = {'name': 'Uniform', 'low': self.kernel_variance_bounds[0],
                                               'high': self.kernel_variance_bounds[1]}
        self.fixed_kernel_variance = (self.kernel_variance_prior_spec == 'fixed')
        if not self.fixed_kernel_variance:
            self.kernel_variance_prior_dist = _pr.make_distribution(self.kernel_variance_prior_spec)
        else:
            self.kernel_variance_prior_dist = _pr.DummyDistribution()

    def _setup_lengthscale_prior(self, lengthscale_prior):
        """BOBE/gp.py:322-337."""
        self.lengthscale_prior_spec = lengthscale_prior
        if self.lengthscale_prior_spec is None:
            self.lengthscale_prior_spec = {'name': 'Uniform', 'low': self.lengthscale_bounds[0],
                                           'high': self.lengthscale_bounds[1]}
        if isinstance(self.lengthscale_prior_spec, str) and self.lengthscale_prior_spec == 'DSLP':
            self.lengthscale_prior_dist = _pr.dslp_distribution(self.ndim)
            self.prior_func = self._standard_prior_logprob
        elif isinstance(self.lengthscale_prior_spec, str) and self.lengthscale_prior_spec == 'SAAS':
            self.lengthscale_prior_dist = None
            self.prior_func = self._saas_prior_logprob
        else:
            self.lengthscale_prior_dist = _pr.make_distribution(self.lengthscale_prior_spec)
            self.prior_func = self._standard_prior_logprob

    @property
    def _is_saas(self):
        return isinstance(self.lengthscale_prior_spec, str) and self.lengthscale_prior_spec == 'SAAS'

    def _setup_optimization_parameters(self):
        """BOBE/gp.py:339-355 -- bounds stored as log(bounds).T, shape (2, P)."""
        self.hyperparam_names = ['lengthscales']
        bounds = [list(self.lengthscale_bounds)] * self.ndim
        if not self.fixed_kernel_variance:
            self.hyperparam_names.append('kernel_variance')
            bounds.append(list(self.kernel_variance_bounds))
        if self._is_saas:
            self.hyperparam_names.append('tausq')
            bounds.append(list(self.tausq_bounds))
        self.hyperparam_bounds = np.log(np.array(bounds, dtype=np.float64).T)
        self.num_hyperparams = self.hyperparam_bounds.shape[1]

    def _standard_prior_logprob(self, lengthscales, kernel_variance, tausq=None):
        """BOBE/gp.py:357-362."""
        logprior = float(np.sum(self.kernel_variance_prior_dist.log_prob(kernel_variance)))
        if self.lengthscale_prior_dist is not None:
            logprior += float(np.sum(self.lengthscale_prior_dist.log_prob(lengthscales)))
        return logprior

    def _saas_prior_logprob(self, lengthscales, kernel_variance, tausq):
        """BOBE/gp.py:364-366."""
        return _pr.saas_prior_logprob(lengthscales, kernel_variance, tausq)

    def _prior_grad(self, lengthscales, kernel_variance, tausq):
        """d log prior / d log_params in the layout of ``_parse_hyperparams``."""
        g = np.zeros(self.num_hyperparams)
        d = self.ndim
        if self._is_saas:
            g_ls, g_kv, g_tau = _pr.saas_prior_grad(lengthscales, kernel_variance, tausq)
            g[:d] = g_ls
            idx = d
            if not self.fixed_kernel_variance:
                g[idx] = g_kv
                idx += 1
            if idx < self.num_hyperparams:
                g[idx] = g_tau
            return g
        if not self.fixed_kernel_variance:
            g[d] = float(np.sum(self.kernel_variance_prior_dist.dlogp_dlogz(kernel_variance)))
        if self.lengthscale_prior_dist is not None:
            g[:d] = self.lengthscale_prior_dist.dlogp_dlogz(lengthscales)
        return g

    def _parse_hyperparams(self, log_params):
        """BOBE/gp.py:368-383."""
        hyperparams = np.exp(np.asarray(log_params, dtype=np.float64))
        lengthscales = hyperparams[:self.ndim]
        if self.fixed_kernel_variance:
            kernel_variance = self.kernel_variance
            if 'tausq' in self.hyperparam_names:
                tausq = hyperparams[self.ndim] if len(hyperparams) > self.ndim else self.tausq
            else:
                tausq = self.tausq
        else:
            kernel_variance = hyperparams[self.ndim]
            tausq = hyperparams[self.ndim + 1] if len(hyperparams) > self.ndim + 1 else self.tausq
        return lengthscales, kernel_variance, tausq

    # ---- device state ----------------------------------------------------------------------------------
    def _invalidate(self):
        self._factor_ok = False
        self._X_dev = self._y_dev = self._ls_dev = self._L_dev = self._Linv_dev = self._alpha_dev = None
        self._LinvT_dev = None
        self._cholesky_np = self._alphas_np = None
        self._info = 0

    @property
    def device(self) -> torch.device:
        return _dev(self._device_arg)

    def _ensure_factor(self):
        if self._factor_ok:
            return
        dev = self.device
        n = self.train_x.shape[0]
        self._X_dev = _to_dev(self.train_x, dev)
        self._y_dev = _to_dev(self.train_y.reshape(-1), dev)
        self._ls_dev = _to_dev(self.lengthscales, dev).reshape(-1)
        if n == 0:
            self._L_dev = torch.zeros((0, 0), dtype=torch.float64, device=dev)
            self._Linv_dev = torch.zeros((0, 0), dtype=torch.float64, device=dev)
            self._alpha_dev = torch.zeros((0,), dtype=torch.float64, device=dev)
            self._factor_ok = True
            return
        kv = torch.tensor([float(self.kernel_variance)], dtype=torch.float64, device=dev)
        L, Linv, alpha, logdet, quad, info = ops.factorize(self.kernel_name, self._X_dev, self._y_dev,
                                                           self._ls_dev[None, :], kv, float(self.noise))
        self._L_dev, self._Linv_dev, self._alpha_dev = L[0], Linv[0], alpha[0]
        self._logdet, self._quad, self._info_dev = logdet, quad, info
        self._factor_hp = (np.array(self.lengthscales, dtype=np.float64).copy(), float(self.kernel_variance),
                           float(self.noise), self.kernel_name)
        self._factor_ok = True

    @property
    def cholesky(self):
        """(n, n) lower factor, zero upper; all-NaN when K is not positive definite (jnp.linalg.cholesky)."""
        if self._cholesky_np is None:
            self._ensure_factor()
            n = self.train_x.shape[0]
            L = self._L_dev[:n, :n].cpu().numpy().copy()
            if n and int(self._info_dev.item()) != 0:
                L[:] = np.nan
            self._cholesky_np = L
        return self._cholesky_np

    @cholesky.setter
    def cholesky(self, value):  # from_state_dict restores the stored factor (BOBE/gp.py:672-673)
        self._cholesky_np = None if value is None else np.array(value, dtype=np.float64)

    @property
    def alphas(self):
        if self._alphas_np is None:
            self._ensure_factor()
            n = self.train_x.shape[0]
            self._alphas_np = self._alpha_dev[:n].cpu().numpy().reshape(-1, 1).copy()
        return self._alphas_np

    @alphas.setter
    def alphas(self, value):
        self._alphas_np = None if value is None else np.array(value, dtype=np.float64).reshape(-1, 1)

    # ---- marginal likelihood -----------------------------------------------------------------------------
    def neg_mll_and_grad_batched(self, log_params):
        """(R, P) log-parameters -> (neg_mll (R,), grad (R, P)); NaN rows where K is not PD.

        One lock-step device call for all rows; prior terms added on the host.
        """
        lp = np.ascontiguousarray(np.atleast_2d(np.asarray(log_params, dtype=np.float64)))
        both_dev = self._mll_grad_device(lp)
        # the device call above is asynchronous: the O(R d) prior terms are evaluated on the host WHILE it runs, and only
        # then are the results fetched (the .cpu() below is the first synchronisation)
        pv, pg = self._prior_terms(lp)
        both = both_dev.cpu().numpy()  # one D2H copy
        return -(both[:, 0] + pv), -(both[:, 1:] + pg)

    def _mll_grad_device(self, lp: np.ndarray) -> torch.Tensor:
        """(R, 1 + P) device tensor [log-ML | gradient] WITHOUT the prior terms, enqueued asynchronously (the sharded
        evaluation all-gathers these rows on the device before anything is fetched: ``dist.mll_grad_sharded``)."""
        self._ensure_factor()
        # (persistent staging buffers: the same device pointers on every optimiser step, so the native call replays its
        # captured CUDA graph instead of enqueueing several hundred launches)
        val, grad, _info = ops.mll_grad_batched(self.kernel_name, self._X_dev, self._y_dev, torch.from_numpy(lp),
                                                not self.fixed_kernel_variance, float(self.kernel_variance),
                                                float(self.noise), reuse_buffers=True)
        return torch.cat([val[:, None], grad], dim=1)

    def _prior_terms(self, lp: np.ndarray):
        """Log-prior values (R,) and gradients (R, P) of the rows of ``lp`` (host)."""
        pv, pg = np.empty(lp.shape[0]), np.empty_like(lp)
        for r in range(lp.shape[0]):
            ls, kv, tausq = self._parse_hyperparams(lp[r])
            pv[r] = self.prior_func(ls, kv, tausq)
            pg[r] = self._prior_grad(ls, kv, tausq)
        return pv, pg

    def neg_mll_and_grad(self, log_params):
        v, g = self.neg_mll_and_grad_batched(np.asarray(log_params, dtype=np.float64)[None, :])
        return float(v[0]), g[0]

    def neg_mll(self, log_params):
        """BOBE/gp.py:385-398."""
        return self.neg_mll_and_grad(log_params)[0]

    def fit(self, x0: np.ndarray = None, maxiter: int = 500) -> dict:
        """BOBE/gp.py:400-437 -- returns {'mll', 'params'}; does NOT apply them (the caller does)."""
        if x0 is None:
            x0 = np.log(self.get_hyperparams())[None, :]
        x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
        optimizer_options = self.optimizer_options.copy()
        kwargs = dict(fun=self.neg_mll, num_params=self.num_hyperparams, bounds=self.hyperparam_bounds, x0=x0,
                      maxiter=maxiter, n_restarts=x0.shape[0], optimizer_options=optimizer_options,
                      value_and_grad=self.neg_mll_and_grad, batched_value_and_grad=self.neg_mll_and_grad_batched)
        if self.mll_optimize is optimize_optax:  # unit-cube coordinates, as the reference's optax path expects
            from .optim import scale_to_unit
            kwargs["x0"] = scale_to_unit(x0, self.hyperparam_bounds)
        best_params_log, best_loss = self.mll_optimize(**kwargs)
        return {'mll': -best_loss, 'params': best_params_log}

    def update_hyperparams(self, hyperparams):
        """BOBE/gp.py:439-448 (argument is in log space, as produced by ``fit``)."""
        lengthscales, kernel_variance, tausq = self._parse_hyperparams(hyperparams)
        self.lengthscales = np.array(lengthscales, dtype=np.float64)
        if not self.fixed_kernel_variance:
            self.kernel_variance = float(kernel_variance)
        self.tausq = float(tausq)
        self.recompute_cholesky()

    def recompute_cholesky(self):
        """BOBE/gp.py:544-550."""
        self._invalidate()
        if torch.cuda.is_available():
            self._ensure_factor()

    # ---- prediction ------------------------------------------------------------------------------------
    _PIPE_ROWS = 8 * 148 * 128  # host-side pipelining granularity: 8 device chunks of bobe_predict

    def _predict(self, x, want_mean, want_var, standardised):
        as_t = _is_t(x)
        self._ensure_factor()
        if self.train_x.shape[0] == 0:
            raise ValueError("GP has no training points")
        if not (as_t and x.is_cuda):
            xh = x if as_t else torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))
            if xh.dim() == 1:
                xh = xh[None, :]
            if xh.shape[1] != self.ndim:
                raise ValueError(f"query points must have {self.ndim} columns")
            if xh.shape[0] > self._PIPE_ROWS:
                mean, var = self._predict_host_pipelined(xh.to(torch.float64), want_mean, want_var, standardised)
                if not as_t:
                    mean = mean.numpy() if mean is not None else None
                    var = var.numpy() if var is not None else None
                return mean, var
        xq = _to_dev(x, self.device)
        if xq.dim() == 1:
            xq = xq[None, :]
        if xq.shape[1] != self.ndim:
            raise ValueError(f"query points must have {self.ndim} columns")
        mean, var = self._predict_dev(xq, want_mean, want_var, standardised)
        if not as_t:
            mean = mean.cpu().numpy() if mean is not None else None
            var = var.cpu().numpy() if var is not None else None
        elif not x.is_cuda:  # CPU tensor in -> CPU tensor out
            mean = mean.cpu() if mean is not None else None
            var = var.cpu() if var is not None else None
        return mean, var

    def _predict_dev(self, xq, want_mean, want_var, standardised):
        return ops.predict(self.kernel_name, self._X_dev, self._ls_dev, float(self.kernel_variance),
                           float(self.noise), self._Linv_dev, self._alpha_dev, xq, float(self.y_mean),
                           float(self.y_std), want_mean, want_var, standardised)

    def _predict_host_pipelined(self, xh, want_mean, want_var, standardised):
        """Large host-resident query sets: the H2D copy of block i+1 (copy stream, double-buffered staging) overlaps
        with the kernels of block i; results return through pinned buffers.  Arithmetic identical to one big call
        (queries are independent)."""
        dev = self.device
        M, d = xh.shape
        rows = self._PIPE_ROWS
        comp = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        stage = [torch.empty((rows, d), dtype=torch.float64, device=dev) for _ in range(2)]
        # the staging blocks come from the caching allocator on the compute stream: a recycled block may still be in use
        # by kernels queued there, so the copy stream must not write into it before those have run
        copy.wait_stream(comp)
        for st in stage:
            st.record_stream(copy)
        copied = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        # straight from torch's caching pinned allocator (no pageable allocation + copy as .pin_memory() would do)
        mean_h = torch.empty(M, dtype=torch.float64, pin_memory=True) if want_mean else None
        var_h = torch.empty(M, dtype=torch.float64, pin_memory=True) if want_var else None
        for i, s in enumerate(range(0, M, rows)):
            e, b = min(M, s + rows), i & 1
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(freed[b])
                stage[b][: e - s].copy_(xh[s:e], non_blocking=True)
                copied[b].record(copy)
            comp.wait_event(copied[b])
            mean, var = self._predict_dev(stage[b][: e - s], want_mean, want_var, standardised)
            freed[b].record(comp)
            if want_mean:
                mean_h[s:e].copy_(mean, non_blocking=True)
            if want_var:
                var_h[s:e].copy_(var, non_blocking=True)
        comp.synchronize()
        return mean_h, var_h

    def predict_mean_single(self, x):
        """BOBE/gp.py:450-457 -- un-standardised mean, scalar."""
        m = self._predict(np.atleast_2d(x) if not _is_t(x) else x, True, False, False)[0]
        return m[0]

    def predict_var_single(self, x):
        """BOBE/gp.py:459-466."""
        v = self._predict(np.atleast_2d(x) if not _is_t(x) else x, False, True, False)[1]
        return v[0]

    def predict_mean_batched(self, x):
        """BOBE/gp.py:468-470."""
        return self._predict(x, True, False, False)[0]

    def predict_var_batched(self, x):
        """BOBE/gp.py:472-474."""
        return self._predict(x, False, True, False)[1]

    def predict_mean_var_batched(self, x):
        """Un-standardised mean and variance in one fused pass (the headline workload)."""
        return self._predict(x, True, True, False)

    def predict_single(self, x):
        """BOBE/gp.py:476-489 -- standardised (mean, var); var has shape (1,)."""
        m, v = self._predict(np.atleast_2d(x) if not _is_t(x) else x, True, True, True)
        return m[0], v[:1]

    def predict_batched(self, x):
        """BOBE/gp.py:491-493 -- ((M,), (M, 1))."""
        m, v = self._predict(x, True, True, True)
        return m, v.reshape(-1, 1)

    # ---- input gradients -----------------------------------------------------------------
    def _ensure_linvT(self):
        """Linv^T on the device, built once per factorisation (the variance gradient needs w = Linv^T (Linv k*))."""
        self._ensure_factor()
        if self._LinvT_dev is None:
            self._LinvT_dev = ops.linv_transpose(self._Linv_dev, self.train_x.shape[0])
        return self._LinvT_dev

    def predict_grad_batched(self, x, standardised=False, want_mean=True, want_var=True):
        """(mean, var, dmean/dx, dvar/dx) at (M, d) points: ``jax.value_and_grad`` of ``predict_mean_single`` /
        ``predict_var_single`` (``standardised=False``) or of ``predict_single`` (``standardised=True``) with respect
        to the query point -- the derivative the reference takes for NUTS on the surrogate
        (``BOBE/samplers.py:268-285``) and for EI / LogEI optimisation (``BOBE/acquisition.py:281-290``).
        Entries not requested are ``None``."""
        as_t = _is_t(x)
        if self.train_x.shape[0] == 0:
            raise ValueError("GP has no training points")
        xq = _to_dev(x, self.device)
        if xq.dim() == 1:
            xq = xq[None, :]
        if xq.shape[1] != self.ndim:
            raise ValueError(f"query points must have {self.ndim} columns")
        self._ensure_factor()
        linvT = self._ensure_linvT() if want_var else None
        out = ops.predict_grad(self.kernel_name, self._X_dev, self._ls_dev, float(self.kernel_variance),
                               float(self.noise), self._Linv_dev, linvT, self._alpha_dev, xq, float(self.y_mean),
                               float(self.y_std), want_mean, want_var, standardised)
        if as_t:
            return out if x.is_cuda else tuple(o.cpu() if o is not None else None for o in out)
        return tuple(o.cpu().numpy() if o is not None else None for o in out)

    def predict_mean_value_and_grad(self, x):
        """``jax.value_and_grad(gp.predict_mean_single)(x)``: un-standardised mean and its (d,) gradient."""
        m, _, dm, _ = self.predict_grad_batched(np.atleast_2d(x) if not _is_t(x) else x, False, True, False)
        return m[0], dm[0]

    def predict_var_value_and_grad(self, x):
        """``jax.value_and_grad(gp.predict_var_single)(x)``."""
        _, v, _, dv = self.predict_grad_batched(np.atleast_2d(x) if not _is_t(x) else x, False, False, True)
        return v[0], dv[0]

    # ---- update ------------------------------------------------------------------------------------------
    def update(self, new_x, new_y):
        """BOBE/gp.py:495-541 -- dedupe, append, re-standardise, re-factor (hyper-parameters unchanged)."""
        new_x = np.atleast_2d(new_x.detach().cpu().numpy() if _is_t(new_x) else np.asarray(new_x, dtype=np.float64))
        new_y = np.atleast_2d(new_y.detach().cpu().numpy() if _is_t(new_y) else np.asarray(new_y, dtype=np.float64))
        new_pts_to_add, new_vals_to_add = [], []
        for i in range(new_x.shape[0]):
            if np.any(np.all(np.isclose(self.train_x, new_x[i], atol=1e-6, rtol=1e-4), axis=1)):
                log.debug(f"Point {new_x[i]} already exists in the training set, not updating")
            else:
                new_pts_to_add.append(new_x[i])
                new_vals_to_add.append(new_y[i])
        if new_pts_to_add:
            new_pts_to_add = np.array(new_pts_to_add)
            new_vals_to_add = np.array(new_vals_to_add).reshape(len(new_vals_to_add), -1)
            self.train_x = np.vstack([self.train_x, new_pts_to_add])
            train_y_original = np.vstack([self.train_y * self.y_std + self.y_mean, new_vals_to_add])
            self.y_mean = float(np.mean(train_y_original))
            self.y_std = float(np.std(train_y_original))
            if self.y_std == 0:
                log.warning("Training targets have zero variance. Setting std to 1.0 to avoid division by zero.")
                self.y_std = 1.0
            self.train_y = (train_y_original - self.y_mean) / self.y_std
            n_old = self.train_x.shape[0] - new_pts_to_add.shape[0]
            if not (self.incremental_update and self._append_factor(n_old)):
                self.recompute_cholesky()

    incremental_update = True  # GP.update extends the factor in O(b n^2) instead of re-factorising

    def _append_factor(self, n_old) -> bool:
        """Rank-b extension of the device factor after ``update`` appended rows to ``train_x`` (hyper-parameters are
        unchanged there, BOBE/gp.py:541).  Returns False (caller re-factorises) when there is no valid factor to
        extend or an appended pivot is not positive."""
        if not (self._factor_ok and n_old > 0 and torch.cuda.is_available()):
            return False
        hp = getattr(self, "_factor_hp", None)  # the factor must belong to the CURRENT hyper-parameters (a caller may have
        if hp is None or not (np.array_equal(hp[0], np.asarray(self.lengthscales, dtype=np.float64).reshape(-1))  # set
                              and hp[1] == float(self.kernel_variance) and hp[2] == float(self.noise)  # the attributes
                              and hp[3] == self.kernel_name):                                           # directly)
            return False
        if int(self._info_dev.reshape(-1)[0].item()) != 0:
            return False
        dev = self.device
        X = _to_dev(self.train_x, dev)
        y = _to_dev(self.train_y.reshape(-1), dev)
        L, Linv, alpha, info = ops.factor_append(self.kernel_name, X, y, n_old, self._ls_dev,
                                                 float(self.kernel_variance), float(self.noise), self._L_dev,
                                                 self._Linv_dev)
        if int(info.item()) != 0:
            return False
        self._X_dev, self._y_dev = X, y
        self._L_dev, self._Linv_dev, self._alpha_dev, self._info_dev = L, Linv, alpha, info
        self._LinvT_dev = None
        self._cholesky_np = self._alphas_np = None
        return True

    # ---- fantasy variance ------------------------------------------------------------------------------
    def fantasy_var(self, new_x, mc_points, k_train_mc=None):
        """BOBE/gp.py:552-576 -- variance at ``mc_points`` if ``new_x`` were added (value-independent).

        ``k_train_mc`` is accepted for interface compatibility and ignored: K(X, MC) is rebuilt on the device
        inside the fused call (it is cheaper than copying it in).  ``new_x`` may hold several candidates
        (C, d); the result is then (C, n_mc).
        """
        as_t = _is_t(new_x) or _is_t(mc_points)
        self._ensure_factor()
        dev = self.device
        cand = _to_dev(new_x, dev)
        single = cand.dim() == 1
        if single:
            cand = cand[None, :]
        out = ops.fantasy_var(self.kernel_name, self._X_dev, self._ls_dev, float(self.kernel_variance),
                              float(self.noise), self._Linv_dev, float(self.y_std), _to_dev(mc_points, dev), cand,
                              "none")
        if single:
            out = out[0]
        return out if as_t else out.cpu().numpy()

    def fantasy_acquisition(self, mc_points, candidates=None, std=False):
        """mean_j fantasy_var (WIPV) or mean_j sqrt(fantasy_var) (WIPStd) for every candidate, one fused call.

        ``candidates=None`` uses the MC points themselves as candidates (BOBE/acquisition.py:390-397).
        """
        as_t = _is_t(mc_points)
        self._ensure_factor()
        dev = self.device
        out = ops.fantasy_var(self.kernel_name, self._X_dev, self._ls_dev, float(self.kernel_variance),
                              float(self.noise), self._Linv_dev, float(self.y_std), _to_dev(mc_points, dev),
                              None if candidates is None else _to_dev(np.atleast_2d(candidates)
                                                                     if not _is_t(candidates) else candidates, dev),
                              "mean_sqrt" if std else "mean")
        return out if as_t else out.cpu().numpy()

    def fantasy_acquisition_value_and_grad(self, mc_points, candidates, std=False):
        """WIPV / WIPStd at the (C, d) candidates and their (C, d) gradients with respect to the candidate point:
        ``jax.value_and_grad(acq.fun)`` as the n <= 500 polish takes it (BOBE/acquisition.py:400-412 through
        BOBE/optim.py:118,309), analytically in one ``bobe_fantasy_var_grad`` call."""
        as_t = _is_t(mc_points) or _is_t(candidates)
        dev = self.device
        cand = _to_dev(candidates, dev)
        if cand.dim() == 1:
            cand = cand[None, :]
        if cand.shape[1] != self.ndim:
            raise ValueError(f"candidates must have {self.ndim} columns")
        linvT = self._ensure_linvT()
        val, grad = ops.fantasy_var_grad(self.kernel_name, self._X_dev, self._ls_dev, float(self.kernel_variance),
                                         float(self.noise), self._Linv_dev, linvT, float(self.y_std),
                                         _to_dev(mc_points, dev), cand, "mean_sqrt" if std else "mean")
        return (val, grad) if as_t else (val.cpu().numpy(), grad.cpu().numpy())

    def get_random_point(self, rng=None, nstd=None):
        """BOBE/gp.py:578-585."""
        rng = rng if rng is not None else np.random.default_rng()
        return rng.uniform(0, 1, size=self.train_x.shape[1])

    # ---- state -------------------------------------------------------------------------------------------
    def state_dict(self):
        """BOBE/gp.py:587-636 -- same keys (pool.gp_fit ships this to workers, bo.py saves it)."""
        return {
            'train_x': np.array(self.train_x),
            'train_y': np.array(self.train_y * self.y_std + self.y_mean),
            'lengthscales': np.array(self.lengthscales),
            'kernel_variance': float(self.kernel_variance),
            'noise': float(self.noise),
            'tausq': float(self.tausq),
            'y_mean': float(self.y_mean),
            'y_std': float(self.y_std),
            'kernel_name': self.kernel_name,
            'lengthscale_prior_spec': self.lengthscale_prior_spec,
            'kernel_variance_prior_spec': self.kernel_variance_prior_spec,
            'fixed_kernel_variance': self.fixed_kernel_variance,
            'optimizer_method': self.optimizer_method,
            'optimizer_options': self.optimizer_options,
            'lengthscale_bounds': self.lengthscale_bounds,
            'kernel_variance_bounds': self.kernel_variance_bounds,
            'tausq_bounds': self.tausq_bounds,
            'cholesky': np.array(self.cholesky) if torch.cuda.is_available() or self._cholesky_np is not None else None,
            'alphas': np.array(self.alphas) if torch.cuda.is_available() or self._alphas_np is not None else None,
            'ndim': self.ndim,
            'gp_class': 'GP',
        }

    @classmethod
    def from_state_dict(cls, state):
        """BOBE/gp.py:638-677."""
        def _plain(v):  # np.load(allow_pickle=True) wraps dicts / lists in 0-d or object arrays
            if isinstance(v, np.ndarray) and v.dtype == object and v.shape == ():
                return v.item()
            if isinstance(v, np.ndarray) and v.dtype.kind in "fiu" and v.ndim == 1 and v.size == 2:
                return v.tolist()
            return v
        gp = cls(train_x=state['train_x'], train_y=state['train_y'], noise=state['noise'],
                 kernel=_plain(state['kernel_name']), optimizer=_plain(state['optimizer_method']),
                 optimizer_options=_plain(state['optimizer_options']), lengthscales=state['lengthscales'],
                 kernel_variance=state['kernel_variance'], lengthscale_bounds=_plain(state['lengthscale_bounds']),
                 kernel_variance_bounds=_plain(state['kernel_variance_bounds']),
                 kernel_variance_prior=_plain(state.get('kernel_variance_prior_spec')),
                 lengthscale_prior=_plain(state.get('lengthscale_prior_spec')), tausq=state.get('tausq', 1.0),
                 tausq_bounds=_plain(state.get('tausq_bounds', [-4, 4])))
        # The reference overwrites cholesky/alphas with the stored arrays (BOBE/gp.py:672-675).  The device
        # state (L^-1, alpha) is always rebuilt from the training data, so predictions never depend on a stale
        # factor; the stored arrays only seed the host-visible attributes.
        if state.get('cholesky') is not None:
            gp.cholesky = np.asarray(state['cholesky'])
        if state.get('alphas') is not None:
            gp.alphas = np.asarray(state['alphas'])
        return gp

    @classmethod
    def load(cls, filename, **kwargs):
        """BOBE/gp.py:679-721."""
        if not filename.endswith('.npz'):
            filename += '.npz'
        try:
            data = np.load(filename, allow_pickle=True)
        except FileNotFoundError:
            raise FileNotFoundError(f"Could not find file {filename}")
        state = {}
        for key in data.files:
            value = data[key]
            state[key] = value.item() if isinstance(value, np.ndarray) and value.shape == () else value
        state.update(kwargs)
        return cls.from_state_dict(state)

    def save(self, filename='gp'):
        """BOBE/gp.py:723-737."""
        if not filename.endswith('.npz'):
            filename += '.npz'
        np.savez(filename, **self.state_dict())

    def copy(self):
        """BOBE/gp.py:740-750."""
        return self.__class__.from_state_dict(self.state_dict())

    @property
    def npoints(self):
        return self.train_x.shape[0]

    def get_hyperparams(self):
        """BOBE/gp.py:756-762."""
        hp = np.asarray(self.lengthscales, dtype=np.float64)
        if not self.fixed_kernel_variance:
            hp = np.hstack([hp, self.kernel_variance])
        if self._is_saas:
            hp = np.hstack([hp, self.tausq])
        return hp

    def hyperparams_dict(self):
        """BOBE/gp.py:764-772."""
        ls_str = {name: f"{float(val):.4f}" for name, val in zip(self.param_names, self.lengthscales)}
        param_dict = {'lengthscales': ls_str, 'kernel_variance': f"{float(self.kernel_variance):.4f}"}
        if 'tausq' in self.hyperparam_names:
            param_dict['tausq'] = f"{float(self.tausq):.4f}"
        return param_dict
