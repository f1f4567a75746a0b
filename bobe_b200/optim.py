"""Multi-restart minimisers with the reference's interface (BOBE/optim.py), JAX-free.

The reference differentiates ``fun`` with ``jax.value_and_grad`` (BOBE/optim.py:118,211,309).  Here the
gradient comes from the caller: ``value_and_grad`` (one point) and/or ``batched_value_and_grad`` (R points
in one GPU call).  With the latter, ``optimize_scipy`` runs the R L-BFGS-B instances in lock step: every
restart keeps scipy's unmodified L-BFGS-B iteration (so its iterates equal the sequential ones), but their
function evaluations are gathered into one batched device call.
"""
from __future__ import annotations

import logging
import threading
from typing import Callable, Optional, Tuple

import numpy as np
from scipy.optimize import minimize

log = logging.getLogger("bobe_b200.optim")


def scale_to_unit(x, param_bounds):
    """BOBE/utils/core.py:181-186."""
    return (x - param_bounds[0]) / (param_bounds[1] - param_bounds[0])


def scale_from_unit(x, param_bounds):
    """BOBE/utils/core.py:188-193."""
    return x * (param_bounds[1] - param_bounds[0]) + param_bounds[0]


def _setup_bounds(bounds, num_params: int) -> Optional[np.ndarray]:
    """BOBE/optim.py:42-69 -- None, shape (2,) shared, or (2, num_params)."""
    if bounds is None:
        return None
    bounds = np.array(bounds, dtype=np.float64)
    if bounds.shape == (2,):
        bounds = np.tile(bounds.reshape(1, 2), (num_params, 1)).T
    elif bounds.shape != (2, num_params):
        raise ValueError(f"Bounds shape {bounds.shape} incompatible with {num_params} parameters")
    return bounds


def _fd_value_and_grad(fun: Callable, h: float = 1e-6) -> Callable:
    """Central differences, for callers that supply neither gradient form."""

    def vg(x):
        x = np.asarray(x, dtype=np.float64)
        f0 = float(fun(x))
        g = np.zeros_like(x)
        for j in range(x.size):
            e = np.zeros_like(x)
            e[j] = h
            g[j] = (float(fun(x + e)) - float(fun(x - e))) / (2 * h)
        return f0, g

    return vg


class LockstepEvaluator:
    """Gathers one pending evaluation per live worker thread into a single batched call.

    The rendezvous costs one short critical section per arrival and one private gate (a ``threading.Lock`` used as a binary
    semaphore) per worker: the LAST arrival evaluates the batch outside the shared lock and opens the other workers' gates,
    so a round of W workers is W uncontended futex wake-ups instead of W re-acquisitions of one condition variable.
    """

    def __init__(self, batched_fn: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray]], n_workers: int):
        self.fn = batched_fn
        self.mu = threading.Lock()
        self.live = n_workers
        self.gates = [threading.Lock() for _ in range(n_workers)]
        for g in self.gates:
            g.acquire()
        self.x = [None] * n_workers
        self.results = [None] * n_workers
        self.pending = []
        self.generation = 0
        self.n_batched_calls = 0
        self.error = None

    def _flush(self, keys, me):
        """Evaluate the pending set (called by exactly one thread, with every other live worker parked on its gate)."""
        keys.sort()
        try:
            vals, grads = self.fn(np.stack([self.x[k] for k in keys]))
            vals, grads = np.asarray(vals, dtype=np.float64), np.asarray(grads, dtype=np.float64)
            for i, k in enumerate(keys):
                self.results[k] = (float(vals[i]), grads[i].copy())
        except Exception as e:  # propagate to every waiting worker
            self.error = e
            for k in keys:
                self.results[k] = e
        self.n_batched_calls += 1
        self.generation += 1
        for k in keys:
            if k != me:
                self.gates[k].release()

    def evaluate(self, wid: int, x: np.ndarray):
        self.x[wid] = np.array(x, dtype=np.float64)
        keys = None
        with self.mu:
            self.pending.append(wid)
            if len(self.pending) == self.live:
                keys, self.pending = self.pending, []
        if keys is not None:
            self._flush(keys, wid)
        else:
            self.gates[wid].acquire()
        r, self.results[wid] = self.results[wid], None
        if isinstance(r, Exception):
            raise r
        return r

    def retire(self, wid: int):
        keys = None
        with self.mu:
            self.live -= 1
            if self.live > 0 and len(self.pending) == self.live:
                keys, self.pending = self.pending, []
        if keys is not None:
            self._flush(keys, -1)


def optimize_scipy(fun: Callable = None, fun_args: Optional[Tuple] = (), fun_kwargs: Optional[dict] = {},
                   num_params: int = 1, bounds=None, x0=None,
                   optimizer_options: Optional[dict] = {"method": "L-BFGS-B", "ftol": 1e-6, "gtol": 1e-6},
                   maxiter: int = 200, n_restarts: int = 4, verbose: bool = False,
                   value_and_grad: Optional[Callable] = None,
                   batched_value_and_grad: Optional[Callable] = None) -> Tuple[np.ndarray, float]:
    """BOBE/optim.py:249-359: evaluate every x0, then L-BFGS-B (``jac=True``) from each; best finite result.

    ``optimizer_options`` is mutated exactly like the reference does (``update`` / ``pop``, :292-294).
    """
    optimizer_options.update({"maxiter": maxiter})
    method = optimizer_options.pop("method", "L-BFGS-B")
    bounds_arr = _setup_bounds(bounds, num_params)
    scipy_bounds = None if bounds_arr is None else [(float(bounds_arr[0, i]), float(bounds_arr[1, i]))
                                                   for i in range(num_params)]
    if x0 is None:
        raise ValueError("x0 must be provided (shape: (n_restarts, num_params) or (num_params,))")
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    if x0.shape[0] < n_restarts:
        raise ValueError(f"x0 provided with {x0.shape[0]} restarts but n_restarts={n_restarts}")
    elif x0.shape[0] > n_restarts:
        x0 = x0[:n_restarts]

    if value_and_grad is None and batched_value_and_grad is None:
        if fun is None:
            raise ValueError("need fun, value_and_grad or batched_value_and_grad")
        value_and_grad = _fd_value_and_grad(lambda x: fun(x, *fun_args, **fun_kwargs))
    elif value_and_grad is not None and (fun_args or fun_kwargs):
        _vg = value_and_grad
        value_and_grad = lambda x: _vg(x, *fun_args, **fun_kwargs)  # noqa: E731

    global_best_f, global_best_params = np.inf, None

    # values at the initial points (BOBE/optim.py:325-333)
    if batched_value_and_grad is not None:
        try:
            vals0, _ = batched_value_and_grad(x0)
            vals0 = np.asarray(vals0, dtype=np.float64)
        except Exception as e:
            log.warning(f"  Initial points: failed with an error: {e}")
            vals0 = np.full(len(x0), np.nan)
    else:
        vals0 = np.full(len(x0), np.nan)
        for i, x_init in enumerate(x0):
            try:
                vals0[i] = float(value_and_grad(x_init)[0])
            except Exception as e:
                log.warning(f"  Initial point {i+1}/{n_restarts}: Failed with an error: {e}")
    for i, val in enumerate(vals0):
        if np.isfinite(val) and val < global_best_f:
            global_best_f, global_best_params = float(val), x0[i].copy()

    results = [None] * len(x0)

    def run_one(i, vg):
        try:
            results[i] = minimize(vg, x0[i], method=method, jac=True, bounds=scipy_bounds, options=dict(optimizer_options))
        except Exception as e:  # BOBE/optim.py:351-354
            if verbose:
                log.warning(f"  Restart {i+1}/{n_restarts}: Failed with an error: {e}")
            results[i] = None

    if batched_value_and_grad is not None and len(x0) > 1:
        # One lock-step group: splitting the restarts into groups that alternate on the device (so that one group's host-side
        # L-BFGS-B bookkeeping hides behind the other's device round) was measured and is SLOWER -- restarts retire early, the
        # average batch at 64 restarts is already ~31, and halving it costs more device efficiency than the overlap returns
        # (profiles/r02/fit_lockstep_groups.txt: 2394 evals/s with one group, 1758 with two).
        ev = LockstepEvaluator(batched_value_and_grad, len(x0))

        def worker(i):
            try:
                run_one(i, lambda x: ev.evaluate(i, x))
            finally:
                ev.retire(i)

        threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(len(x0))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        optimize_scipy.last_batched_calls = ev.n_batched_calls
    else:
        if value_and_grad is None:
            def value_and_grad(x):  # single restart through the batched form
                v, g = batched_value_and_grad(np.atleast_2d(x))
                return float(np.asarray(v)[0]), np.asarray(g)[0]
        for i in range(len(x0)):
            run_one(i, value_and_grad)

    for i, result in enumerate(results):  # BOBE/optim.py:339-349
        if result is None:
            continue
        msg = result.message if isinstance(result.message, str) else result.message.decode()
        is_acceptable = result.success or "ITERATIONS REACHED LIMIT" in msg.upper()
        if is_acceptable and np.isfinite(result.fun) and result.fun < global_best_f:
            global_best_f, global_best_params = float(result.fun), np.array(result.x)
    return np.array(global_best_params), float(global_best_f)


class _Adam:
    """optax.adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0) on arrays with a leading batch axis."""

    def __init__(self, shape, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.m = np.zeros(shape)
        self.v = np.zeros(shape)
        self.t = 0

    def step(self, params, grad):
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * grad
        self.v = self.b2 * self.v + (1 - self.b2) * grad * grad
        mhat = self.m / (1 - self.b1**self.t)
        vhat = self.v / (1 - self.b2**self.t)
        return params - self.lr * mhat / (np.sqrt(vhat) + self.eps)


def _optax_common(optimizer_options):
    early_stop_patience = optimizer_options.pop("early_stop_patience", 25)  # BOBE/optim.py:109-111
    lr = optimizer_options.pop("lr", 1e-3)
    name = optimizer_options.pop("name", "adam")
    if name.lower() != "adam":
        raise ValueError(f"Optimizer '{name}' not available (optax is not installed; Adam is built in)")
    lr = optimizer_options.pop("learning_rate", lr)
    return early_stop_patience, lr


def optimize_optax(fun: Callable = None, fun_args=(), fun_kwargs={}, num_params: int = 1, bounds=None, x0=None,
                   optimizer_options: Optional[dict] = {"name": "adam", "lr": 1e-3, "early_stop_patience": 25},
                   maxiter: int = 200, n_restarts: int = 1, verbose: bool = False,
                   value_and_grad: Optional[Callable] = None,
                   batched_value_and_grad: Optional[Callable] = None) -> Tuple[np.ndarray, float]:
    """BOBE/optim.py:71-163: sequential Adam per restart in unit-cube coordinates, clip to [0,1], patience stop.

    ``x0`` is in unit-cube coordinates when bounds are given, exactly as in the reference (:105,:128-133).
    """
    if x0 is None:
        raise ValueError("x0 must be provided (shape: (n_restarts, num_params))")
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    if x0.shape[0] < n_restarts:
        raise ValueError(f"x0 provided with {x0.shape[0]} restarts but n_restarts={n_restarts}")
    x0 = x0[:n_restarts]
    bounds_arr = _setup_bounds(bounds, num_params)
    early_stop_patience, lr = _optax_common(optimizer_options)
    if value_and_grad is None and batched_value_and_grad is not None:
        def value_and_grad(x):
            v, g = batched_value_and_grad(np.atleast_2d(x))
            return float(np.asarray(v)[0]), np.asarray(g)[0]
    if value_and_grad is None:
        value_and_grad = _fd_value_and_grad(lambda x: fun(x, *fun_args, **fun_kwargs))
    elif fun_args or fun_kwargs:
        _vg = value_and_grad
        value_and_grad = lambda x: _vg(x, *fun_args, **fun_kwargs)  # noqa: E731
    if bounds_arr is not None:
        span = bounds_arr[1] - bounds_arr[0]

        def scaled_vg(u):
            v, g = value_and_grad(scale_from_unit(u, bounds_arr))
            return v, np.asarray(g) * span
    else:
        scaled_vg = value_and_grad

    if batched_value_and_grad is not None and not (fun_args or fun_kwargs):
        return _optax_lockstep(batched_value_and_grad, x0, bounds_arr, lr, early_stop_patience, maxiter)

    global_best_f, global_best_params = np.inf, None
    for x_init in x0:  # BOBE/optim.py:128-136
        try:
            val = scaled_vg(x_init)[0]
            if np.isfinite(val) and val < global_best_f:
                global_best_f, global_best_params = float(val), x_init.copy()
        except Exception as e:
            log.warning(f"  Initial point failed with an error: {e}")
    for r in range(n_restarts):  # BOBE/optim.py:138-160
        params = x0[r].copy()
        opt = _Adam(params.shape, lr)
        best_f, patience = float("inf"), early_stop_patience
        for _ in range(maxiter):
            val, grad = scaled_vg(params)
            params = opt.step(params, np.asarray(grad))
            if bounds_arr is not None:
                params = np.clip(params, 0.0, 1.0)
            if val < best_f:
                best_f, patience = float(val), early_stop_patience
            else:
                patience -= 1
                if patience == 0:
                    break
        if best_f < global_best_f:
            global_best_f, global_best_params = best_f, params
    best = scale_from_unit(global_best_params, bounds_arr) if bounds_arr is not None else global_best_params
    return np.array(best), float(global_best_f)


def _optax_lockstep(batched_value_and_grad, x0, bounds_arr, lr, early_stop_patience, maxiter):
    """The sequential per-restart Adam loops of BOBE/optim.py:128-160 advanced in lock step: the restarts never
    interact, so stepping all still-active ones with one batched evaluation gives exactly the trajectories of the
    sequential loops (per-restart patience, per-restart stop, the winner's LAST iterate is returned, :158-160) with
    R times fewer device calls."""
    R = x0.shape[0]
    span = (bounds_arr[1] - bounds_arr[0]) if bounds_arr is not None else 1.0

    def scaled(us):
        xs = scale_from_unit(us, bounds_arr) if bounds_arr is not None else us
        v, g = batched_value_and_grad(xs)
        return np.asarray(v, dtype=np.float64), np.asarray(g, dtype=np.float64) * span

    global_best_f, global_best_params = np.inf, None
    try:  # initial sweep over x0 (:128-136)
        v0, _ = scaled(x0)
        for r in range(R):
            if np.isfinite(v0[r]) and v0[r] < global_best_f:
                global_best_f, global_best_params = float(v0[r]), x0[r].copy()
    except Exception as e:
        log.warning(f"  Initial points failed with an error: {e}")
    params = x0.copy()
    m, v = np.zeros_like(params), np.zeros_like(params)
    best_f = np.full(R, np.inf)
    patience = np.full(R, early_stop_patience, dtype=np.int64)
    active = np.ones(R, dtype=bool)
    b1, b2, eps = 0.9, 0.999, 1e-8
    for t in range(1, maxiter + 1):
        idx = np.where(active)[0]
        if idx.size == 0:
            break
        vals, grads = scaled(params[idx])
        m[idx] = b1 * m[idx] + (1 - b1) * grads
        v[idx] = b2 * v[idx] + (1 - b2) * grads * grads
        step = lr * (m[idx] / (1 - b1**t)) / (np.sqrt(v[idx] / (1 - b2**t)) + eps)
        new = params[idx] - step
        params[idx] = np.clip(new, 0.0, 1.0) if bounds_arr is not None else new
        improved = vals < best_f[idx]
        best_f[idx] = np.where(improved, vals, best_f[idx])
        patience[idx] = np.where(improved, early_stop_patience, patience[idx] - 1)
        active[idx] = patience[idx] != 0
    for r in range(R):  # same winner rule and order as the sequential loop
        if best_f[r] < global_best_f:
            global_best_f, global_best_params = float(best_f[r]), params[r]
    best = scale_from_unit(global_best_params, bounds_arr) if bounds_arr is not None else global_best_params
    return np.array(best), float(global_best_f)


def optimize_optax_vmap(fun: Callable = None, fun_args=(), fun_kwargs={}, num_params: int = 1, bounds=None, x0=None,
                        optimizer_options: Optional[dict] = {"name": "adam", "lr": 1e-3, "early_stop_patience": 25},
                        maxiter: int = 200, n_restarts: int = 1, verbose: bool = False,
                        batched_value_and_grad: Optional[Callable] = None,
                        value_and_grad: Optional[Callable] = None) -> Tuple[np.ndarray, float]:
    """BOBE/optim.py:166-246: all restarts advance in lock step (the reference vmaps; here one batched call)."""
    if x0 is None:
        raise ValueError("x0 must be provided (shape: (n_restarts, num_params))")
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    if x0.shape[0] != n_restarts:
        raise ValueError(f"x0 has {x0.shape[0]} restarts but n_restarts was set to {n_restarts}")
    bounds_arr = _setup_bounds(bounds, num_params)
    early_stop_patience, lr = _optax_common(optimizer_options)
    if batched_value_and_grad is None:
        if value_and_grad is None:
            value_and_grad = _fd_value_and_grad(lambda x: fun(x, *fun_args, **fun_kwargs))

        def batched_value_and_grad(xs):
            out = [value_and_grad(x) for x in xs]
            return np.array([o[0] for o in out]), np.stack([o[1] for o in out])
    span = (bounds_arr[1] - bounds_arr[0]) if bounds_arr is not None else 1.0

    def scaled(us):
        xs = scale_from_unit(us, bounds_arr) if bounds_arr is not None else us
        v, g = batched_value_and_grad(xs)
        return np.asarray(v, dtype=np.float64), np.asarray(g, dtype=np.float64) * span

    params = x0.copy()
    opt = _Adam(params.shape, lr)
    best_vals = np.full(n_restarts, np.inf)
    best_params = np.zeros_like(params)
    patience = np.full(n_restarts, early_stop_patience, dtype=np.int64)
    for _ in range(maxiter):
        vals, grads = scaled(params)
        params = opt.step(params, grads)
        if bounds_arr is not None:
            params = np.clip(params, 0.0, 1.0)
        improved = vals < best_vals
        best_vals = np.where(improved, vals, best_vals)
        best_params = np.where(improved[:, None], params, best_params)
        patience = np.where(improved, early_stop_patience, patience - 1)
        if np.all(patience <= 0):
            break
    i = int(np.argmin(best_vals))
    best = scale_from_unit(best_params[i], bounds_arr) if bounds_arr is not None else best_params[i]
    return np.array(best), float(best_vals[i])
