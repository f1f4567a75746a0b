"""Multi-GPU sharding of the GP hot path: one process per GPU (``torch.distributed``), replicated training-set
state, sharded queries / restarts / candidates, and a closing all-gather of the small results.

This replaces the reference's MPI restart farm (BOBE/pool.py:239-328: ``np.array_split(x0, size)`` + pickled
send/recv + ``max(mll)``) and gives the nested-sampling / WIPV sweeps (BOBE/samplers.py:172,
BOBE/acquisition.py:390-394) a query-sharded form.  No data-path collective is needed: every rank factorises
the same K redundantly (bit-identical replicas, no traffic - SURVEY.md 8e), so the only exchange is the
gather of (value, params) per restart (KB) or of the sharded outputs (16 B/query).

Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU (host-logic tests).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np
import torch
import torch.distributed as tdist


def world() -> Tuple[int, int]:
    if tdist.is_available() and tdist.is_initialized():
        return tdist.get_rank(), tdist.get_world_size()
    return 0, 1


def shard_bounds(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block partition, sizes differing by at most one (same split as ``np.array_split``)."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _comm_device() -> torch.device:
    if tdist.is_initialized() and tdist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allgather_rows(local: torch.Tensor, total: int) -> torch.Tensor:
    """Concatenate row-sharded tensors (shards from ``shard_bounds``) from all ranks, in rank order."""
    rank, ws = world()
    if ws == 1:
        return local
    dev = _comm_device()
    sizes = [shard_bounds(total, r, ws)[1] - shard_bounds(total, r, ws)[0] for r in range(ws)]
    maxrows = max(sizes)
    tail = tuple(local.shape[1:])
    pad = torch.zeros((maxrows,) + tail, dtype=local.dtype, device=dev)
    pad[: local.shape[0]] = local.to(dev)
    out = [torch.empty_like(pad) for _ in range(ws)]
    tdist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0).to(local.device)


def predict_sharded(gp, xq, want_var: bool = True, gather: bool = True):
    """Query-sharded posterior mean (+ variance).  ``xq`` is the FULL (M, d) query set on every rank (host array
    or tensor); each rank evaluates its block and, if ``gather``, all ranks receive the full result."""
    rank, ws = world()
    M = xq.shape[0]
    lo, hi = shard_bounds(M, rank, ws)
    mean, var = gp._predict(xq[lo:hi], True, want_var, False)
    if not gather or ws == 1:
        return mean, var
    as_np = not isinstance(mean, torch.Tensor)
    mt = torch.as_tensor(mean)
    mean_all = allgather_rows(mt, M)
    var_all = allgather_rows(torch.as_tensor(var), M) if want_var else None
    if as_np:
        return mean_all.cpu().numpy(), (var_all.cpu().numpy() if want_var else None)
    return mean_all, var_all


def fit_sharded(fit_fn: Callable[[np.ndarray], dict], x0: np.ndarray) -> dict:
    """Restart-sharded hyper-parameter fit: rank r optimises ``np.array_split(x0, world)[r]`` with ``fit_fn``
    (e.g. ``lambda chunk: gp.fit(chunk, maxiter)``), then every rank receives the best {'mll','params'}.
    Mirrors BOBE/pool.py:299-326 with an all-gather in place of pickled send/recv."""
    rank, ws = world()
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    lo, hi = shard_bounds(x0.shape[0], rank, ws)
    P = x0.shape[1]
    if hi > lo:
        res = fit_fn(x0[lo:hi])
        mll = float(res['mll'])
        params = np.asarray(res['params'], dtype=np.float64).reshape(-1)
        if params.shape != (P,) or not np.isfinite(mll):
            mll, params = -np.inf, np.full(P, np.nan)
    else:
        mll, params = -np.inf, np.full(P, np.nan)
    if ws == 1:
        return {'mll': mll, 'params': params}
    dev = _comm_device()
    mine = torch.tensor(np.concatenate([[mll], params]), dtype=torch.float64, device=dev)
    out = [torch.empty_like(mine) for _ in range(ws)]
    tdist.all_gather(out, mine)
    allr = torch.stack(out).cpu().numpy()
    best = int(np.argmax(allr[:, 0]))  # BOBE/pool.py:324: max over the ranks' best mll
    return {'mll': float(allr[best, 0]), 'params': allr[best, 1:].copy()}


def mll_grad_sharded(gp, log_params: np.ndarray):
    """Restart-sharded lock-step evaluation of (neg_mll, grad) for an (R, P) block; all ranks get all rows."""
    rank, ws = world()
    lp = np.atleast_2d(np.asarray(log_params, dtype=np.float64))
    R = lp.shape[0]
    if ws > 1 and _comm_device().type == "cuda" and hasattr(gp, "_mll_grad_device"):
        return _mll_grad_sharded_device(gp, lp, rank, ws)
    lo, hi = shard_bounds(R, rank, ws)
    if hi > lo:
        v, g = gp.neg_mll_and_grad_batched(lp[lo:hi])
    else:
        v, g = np.zeros(0), np.zeros((0, lp.shape[1]))
    if ws == 1:
        return v, g
    both = torch.as_tensor(np.concatenate([v[:, None], g], axis=1))
    allr = allgather_rows(both, R).cpu().numpy()
    return allr[:, 0], allr[:, 1:]


_gather_buffers: dict = {}


def _mll_grad_sharded_device(gp, lp: np.ndarray, rank: int, ws: int):
    """NCCL route of :func:`mll_grad_sharded`: the shard's [log-ML | gradient] rows never visit the host before the
    collective -- device call (asynchronous), ONE ``all_gather_into_tensor`` on persistent buffers behind it, the prior terms
    of all R rows on the host meanwhile, then a single device-to-host copy (the only synchronisation of the round)."""
    R, P = lp.shape
    sizes = [shard_bounds(R, r, ws)[1] - shard_bounds(R, r, ws)[0] for r in range(ws)]
    maxrows = max(sizes)
    lo, hi = shard_bounds(R, rank, ws)
    dev = _comm_device()
    key = (maxrows, P, ws, dev)
    if key not in _gather_buffers:
        _gather_buffers[key] = (torch.zeros((maxrows, P + 1), dtype=torch.float64, device=dev),
                                torch.empty((ws * maxrows, P + 1), dtype=torch.float64, device=dev))
    mine, out = _gather_buffers[key]
    if hi > lo:
        mine[: hi - lo].copy_(gp._mll_grad_device(np.ascontiguousarray(lp[lo:hi])))
    tdist.all_gather_into_tensor(out, mine)
    pv, pg = gp._prior_terms(lp)
    allr = out.cpu().numpy().reshape(ws, maxrows, P + 1)
    both = np.concatenate([allr[r, : sizes[r]] for r in range(ws)], axis=0)
    return -(both[:, 0] + pv), -(both[:, 1:] + pg)


def acquisition_sharded(eval_fn: Callable[[np.ndarray], np.ndarray], candidates: np.ndarray) -> np.ndarray:
    """Candidate-sharded acquisition sweep (e.g. WIPV over candidates): every rank returns all values."""
    rank, ws = world()
    C = candidates.shape[0]
    lo, hi = shard_bounds(C, rank, ws)
    vals = np.asarray(eval_fn(candidates[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
    if ws == 1:
        return vals
    return allgather_rows(torch.as_tensor(vals), C).cpu().numpy()


def wipv_sharded(gp, mc_points, candidates=None, std: bool = False):
    """WIPV / WIPStd with the MONTE-CARLO COLUMNS sharded (BASELINE config 5: n_mc = 1e5, C = 8).

    The reference maps the candidates over ONE shared ``k_train_mc`` (BOBE/acquisition.py:385-394) and averages the
    fantasy variance over the MC points (BOBE/gp.py:552-576, acquisition.py:438-440,463-465).  The dominant cost is the
    shared solve ``V = L^-1 K(X, MC)`` (n^2 n_mc flops), so a candidate split would make every rank redo all of it;
    here rank r solves only its ``n_mc / G`` columns for ALL candidates and contributes the partial mean
    ``(n_r / n_mc) mean_{j in shard r} s_j(c)``.  The C partial means are all-gathered and added in rank order on every
    rank (fixed order: every rank returns bit-identical values; they differ from the single-GPU value only by the
    regrouping of the mean, ~1e-16 relative).

    ``candidates=None`` uses the full MC set as candidates (acquisition.py:390-397).  Returns (C,) values."""
    rank, ws = world()
    as_t = isinstance(mc_points, torch.Tensor)
    n_mc = int(mc_points.shape[0])
    cand = mc_points if candidates is None else candidates
    C = int(cand.shape[0]) if getattr(cand, "ndim", 2) > 1 else 1
    if ws == 1:
        return gp.fantasy_acquisition(mc_points, candidates, std=std)
    lo, hi = shard_bounds(n_mc, rank, ws)
    if hi > lo:
        part = torch.as_tensor(gp.fantasy_acquisition(mc_points[lo:hi], cand, std=std)).to(torch.float64).reshape(-1)
        part = part * (float(hi - lo) / float(n_mc))
    else:
        part = torch.zeros(C, dtype=torch.float64)
    dev = _comm_device()
    mine = part.to(dev)
    out = [torch.empty_like(mine) for _ in range(ws)]
    tdist.all_gather(out, mine)
    total = out[0].clone()
    for r in range(1, ws):  # rank order, on every rank
        total += out[r]
    total = total.to(part.device)
    return total if as_t else total.cpu().numpy()
