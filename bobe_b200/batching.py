"""Batching adapters for the reference's one-point-at-a-time callers.

The reference's samplers call the surrogate one point at a time: dynesty evaluates ``loglike(x)`` inside each
proposal walk (``BOBE/samplers.py:111-160``), and the initial live points / the posterior-variance pass go through
``jax.lax.map(f, xs, batch_size=...)`` (``samplers.py:131,153,172``).  On the GPU a single-point call is pure launch
latency (~20 us for ~1e5 flops of work), so the throughput of ``bobe_predict`` only reaches the sampler if the
points arrive in batches.  Two adapters, both host-side only (the device work is ``GP.predict_*_batched``):

* :func:`lax_map` -- drop-in for ``jax.lax.map(f, xs, batch_size=b)`` when ``f`` is one of the GP's single-point
  predictors: one batched call instead of ``len(xs)`` dispatches.
* :class:`SurrogatePool` -- a ``pool`` object for ``dynesty`` (``pool.map`` + ``pool.size``, used with
  ``queue_size=pool.size``).  ``map`` runs the mapped tasks (dynesty's proposal walks) on threads; every call the
  tasks make to ``pool.loglike`` blocks until each live task has one point pending, and the whole set goes to the
  device as ONE ``predict_mean_batched`` call.  The walks advance in lock step, ``queue_size`` points per launch.

The same rendezvous (``optim.LockstepEvaluator``) already drives the multi-restart L-BFGS-B fit.
"""
from __future__ import annotations

import threading
from typing import Callable, Iterable, List, Optional

import numpy as np

from .optim import LockstepEvaluator


def lax_map(gp, method: str, xs, batch_size: Optional[int] = None) -> np.ndarray:
    """``jax.lax.map(getattr(gp, method), xs, batch_size=...)`` as one batched device call.

    ``method`` is the name of a single-point predictor (``predict_mean_single``, ``predict_var_single``,
    ``predict_single``); ``batch_size`` is accepted for signature compatibility (the device path chunks internally).
    """
    batched = {"predict_mean_single": "predict_mean_batched", "predict_var_single": "predict_var_batched",
               "predict_single": "predict_batched"}
    if method not in batched:
        raise ValueError(f"lax_map: {method!r} is not a single-point GP predictor")
    xs = np.atleast_2d(np.asarray(xs, dtype=np.float64))
    if xs.shape[0] == 0:
        return np.zeros((0,), dtype=np.float64)
    return getattr(gp, batched[method])(xs)


class SurrogatePool:
    """dynesty-compatible pool that turns concurrent single-point surrogate calls into batched device calls.

    Usage (mirrors ``BOBE/samplers.py:157-160``)::

        pool = SurrogatePool(gp, size=64)
        sampler = StaticNestedSampler(pool.loglike, prior_transform, ndim, pool=pool, queue_size=pool.size, ...)

    ``fn_batched`` overrides what is evaluated: a callable ``(m, d) -> (m,)`` (default: the GP's posterior mean, the
    log-likelihood surrogate of ``samplers.py:111-114``).
    """

    def __init__(self, gp=None, size: int = 32, fn_batched: Optional[Callable[[np.ndarray], np.ndarray]] = None):
        if size < 1:
            raise ValueError("size must be >= 1")
        if fn_batched is None:
            if gp is None:
                raise ValueError("either gp or fn_batched is required")
            fn_batched = gp.predict_mean_batched
        self.size = int(size)
        self._fn = fn_batched
        self._tls = threading.local()
        self._ev: Optional[LockstepEvaluator] = None
        self._threads: List[threading.Thread] = []
        self._go: List[threading.Lock] = []
        self._job = None
        self._mu = threading.Lock()
        self._done = threading.Lock()
        self._done.acquire()
        self.n_device_calls = 0
        self.n_points = 0

    # ---- evaluation -------------------------------------------------------------------------------------------
    def _batched(self, xs: np.ndarray):
        vals = np.asarray(self._fn(xs), dtype=np.float64).reshape(-1)
        self.n_device_calls += 1
        self.n_points += xs.shape[0]
        return vals, np.zeros((xs.shape[0], 0))

    def loglike(self, x) -> float:
        """Single-point surrogate value.  Inside ``map`` it joins the current batch; outside it is a batch of one."""
        wid = getattr(self._tls, "wid", None)
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        if wid is None or self._ev is None:
            return float(self._batched(x[None, :])[0][0])
        return self._ev.evaluate(wid, x)[0]

    __call__ = loglike

    # ---- pool protocol ----------------------------------------------------------------------------------------
    def _worker(self, i: int):
        """Persistent task thread ``i`` (started on the first ``map``; 64 thread starts cost more than a whole round)."""
        self._tls.wid = None
        go = self._go[i]
        while True:
            go.acquire()
            job = self._job
            if job is None:
                return
            fn, items, s, out, errors, ev, state = job
            self._tls.wid = i
            try:
                out[s + i] = fn(items[s + i])
            except BaseException as e:  # noqa: BLE001 -- re-raised on the caller's thread by map()
                errors.append(e)
            finally:
                self._tls.wid = None
                ev.retire(i)
                with self._mu:
                    state[0] -= 1
                    last = state[0] == 0
                if last:
                    self._done.release()

    def _ensure_threads(self, count: int):
        while len(self._threads) < count:
            i = len(self._threads)
            g = threading.Lock()
            g.acquire()
            self._go.append(g)
            t = threading.Thread(target=self._worker, args=(i,), daemon=True, name=f"bobe-surrogate-{i}")
            self._threads.append(t)
            t.start()

    def map(self, fn: Callable, iterable: Iterable) -> List:
        """Run ``fn(item)`` for every item, ``size`` at a time on threads whose ``loglike`` calls are fused."""
        items = list(iterable)
        out: List = [None] * len(items)
        for s in range(0, len(items), self.size):
            count = min(self.size, len(items) - s)
            self._ensure_threads(count)
            ev = LockstepEvaluator(self._batched, count)
            errors: List[BaseException] = []
            self._ev = ev
            self._job = (fn, items, s, out, errors, ev, [count])
            for i in range(count):
                self._go[i].release()
            self._done.acquire()
            self._ev = None
            if errors:
                raise errors[0]
        return out

    def close(self):
        """Stop the task threads (the pool can be used again afterwards: they restart on the next ``map``)."""
        self._job = None
        for g in self._go:
            g.release()
        for t in self._threads:
            t.join(timeout=1.0)
        self._threads, self._go = [], []

    def join(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
