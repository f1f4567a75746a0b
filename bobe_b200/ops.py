"""Tensor-level wrappers over the C-ABI: torch supplies device memory and the stream, nothing else.

Every function takes/returns CUDA float64 tensors and launches on ``torch.cuda.current_stream()``.
No function here has a CPU path; a CPU tensor raises.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import lib, check

KIND = {"rbf": _lib.KERNEL_RBF, "matern": _lib.KERNEL_MATERN52}

_ws_cache = OrderedDict()
_WS_MAX_ARENAS = 4  # most recently used (device, stream) arenas kept alive


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch arena per (device, stream) (the C-ABI never allocates).

    Calls on one stream are stream-ordered, so they can share an arena; calls on different streams (or from different
    host threads with their own current stream) must not."""
    dev = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        _ws_cache[key] = None  # drop the old arena first (the caching allocator keeps it valid for queued kernels)
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=f"cuda:{dev}")
        _ws_cache[key] = buf
    _ws_cache.move_to_end(key)
    while len(_ws_cache) > _WS_MAX_ARENAS:  # arenas of streams no longer in use (stream-ordered free: safe)
        _ws_cache.popitem(last=False)
    return buf


def release_workspace():
    _ws_cache.clear()
    _mll_io_cache.clear()


def _chk(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float64:
        raise TypeError(f"{name} must be a CUDA float64 tensor (no CPU fallback)")
    return t.contiguous()


def npad(n: int) -> int:
    return int(lib.bobe_npad(int(n)))


def kernel_matrix(kind: str, xa, xb, ls, kv: float, noise: float, add_noise: bool) -> torch.Tensor:
    """K(xa, xb) [+ noise I] -- BOBE/gp.py:124-168."""
    xa, xb, ls = _chk(xa, "xa"), _chk(xb, "xb"), _chk(ls, "ls")
    n1, d = xa.shape
    n2 = xb.shape[0]
    out = torch.empty((n1, n2), dtype=torch.float64, device=xa.device)
    with torch.cuda.device(xa.device):
        check(lib.bobe_kernel_matrix(_stream(), KIND[kind], xa.data_ptr(), n1, xb.data_ptr(), n2, d, ls.data_ptr(),
                                     float(kv), float(noise), int(bool(add_noise)), out.data_ptr(), n2),
              "bobe_kernel_matrix")
    return out


def factorize(kind: str, X, y, ls, kv, noise: float, want_L: bool = True):
    """Batched K -> (L, Linv, alpha, logdet, quad, info).  ls (B,d), kv (B,).  BOBE/gp.py:258-260,544-550."""
    X, y, ls, kv = _chk(X, "X"), _chk(y, "y").reshape(-1), _chk(ls, "ls"), _chk(kv, "kv")
    if ls.dim() == 1:
        ls = ls[None, :]
    kv = kv.reshape(-1)
    B = ls.shape[0]
    n, d = X.shape
    p = npad(n)
    dev = X.device
    L = torch.empty((B, p, p), dtype=torch.float64, device=dev) if want_L else None
    Linv = torch.empty((B, p, p), dtype=torch.float64, device=dev)
    alpha = torch.empty((B, p), dtype=torch.float64, device=dev)
    logdet = torch.empty(B, dtype=torch.float64, device=dev)
    quad = torch.empty(B, dtype=torch.float64, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib.bobe_factorize_workspace_bytes(n, d, B)
        ws = _workspace(nbytes, dev)
        check(lib.bobe_factorize(_stream(), KIND[kind], X.data_ptr(), y.data_ptr(), n, d, ls.data_ptr(), kv.data_ptr(),
                                 float(noise), B, L.data_ptr() if want_L else None, Linv.data_ptr(), alpha.data_ptr(),
                                 logdet.data_ptr(), quad.data_ptr(), info.data_ptr(), ws.data_ptr(), ws.numel()),
              "bobe_factorize")
    return L, Linv, alpha, logdet, quad, info


def cholesky_solve(K, y):
    """Factorise caller-supplied kernel matrices K (B, n, n) (or (n, n)) and solve for ``y`` (n,):
    returns (L (B, npad, npad), alpha (B, npad), logdet (B), quad (B), info (B)) -- ``jnp.linalg.cholesky`` +
    ``cho_solve`` of gp_mll, BOBE/gp.py:170-178.  Only the lower triangle of K is read."""
    K, y = _chk(K, "K"), _chk(y, "y").reshape(-1)
    if K.dim() == 2:
        K = K[None]
    B, n, n2 = K.shape
    if n != n2 or y.shape[0] != n:
        raise ValueError("cholesky_solve: K must be (B, n, n) and y (n,)")
    p = npad(n)
    dev = K.device
    L = torch.empty((B, p, p), dtype=torch.float64, device=dev)
    alpha = torch.empty((B, p), dtype=torch.float64, device=dev)
    logdet = torch.empty(B, dtype=torch.float64, device=dev)
    quad = torch.empty(B, dtype=torch.float64, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.bobe_cholesky_workspace_bytes(n, B), dev)
        check(lib.bobe_cholesky_batched(_stream(), K.data_ptr(), n, n, B, y.data_ptr(), L.data_ptr(), None, alpha.data_ptr(),
                                        logdet.data_ptr(), quad.data_ptr(), info.data_ptr(), ws.data_ptr(), ws.numel()),
              "bobe_cholesky_batched")
    return L, alpha, logdet, quad, info


def dist_sq(xa, xb) -> torch.Tensor:
    """sum_k (xa_ik - xb_jk)^2 by direct differences -- BOBE/gp.py:80-96."""
    xa, xb = _chk(xa, "xa"), _chk(xb, "xb")
    n1, d = xa.shape
    n2 = xb.shape[0]
    if xb.shape[1] != d:
        raise ValueError("dist_sq: column counts differ")
    out = torch.empty((n1, n2), dtype=torch.float64, device=xa.device)
    with torch.cuda.device(xa.device):
        check(lib.bobe_dist_sq(_stream(), xa.data_ptr(), n1, xb.data_ptr(), n2, d, out.data_ptr(), n2), "bobe_dist_sq")
    return out


def factor_append(kind: str, X, y, n_old: int, ls, kv: float, noise: float, L, Linv):
    """Extend the padded factors of the first ``n_old`` rows of X by the remaining rows, in O(b n^2), and re-solve
    alpha for all targets ``y`` -- GP.update (BOBE/gp.py:495-541) without the full re-factorisation.

    ``L`` / ``Linv`` are the padded (npad(n_old), npad(n_old)) factors; returns ``(L, Linv, alpha, info)`` padded to
    npad(n) -- the same buffers, extended in place, when the padded size does not change."""
    X, y, ls, L, Linv = _chk(X, "X"), _chk(y, "y").reshape(-1), _chk(ls, "ls"), _chk(L, "L"), _chk(Linv, "Linv")
    n, d = X.shape
    b = n - n_old
    if b <= 0:
        raise ValueError("factor_append: nothing to append")
    p_new, p_old = npad(n), L.shape[0]
    dev = X.device
    if p_new != p_old:  # grow: identity-padded copies (O(n^2) traffic, still no O(n^3) work)
        def grow(A):
            out = torch.zeros((p_new, p_new), dtype=torch.float64, device=dev)
            out[:p_old, :p_old] = A
            idx = torch.arange(p_old, p_new, device=dev)
            out[idx, idx] = 1.0
            return out
        L, Linv = grow(L), grow(Linv)
    alpha = torch.empty(p_new, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.bobe_factor_append_workspace_bytes(n, d), dev)
        check(lib.bobe_factor_append(_stream(), KIND[kind], X.data_ptr(), y.data_ptr(), n_old, b, d, ls.data_ptr(),
                                     float(kv), float(noise), L.data_ptr(), Linv.data_ptr(), alpha.data_ptr(),
                                     info.data_ptr(), ws.data_ptr(), ws.numel()), "bobe_factor_append")
    return L, Linv, alpha, info


_mll_io_cache = OrderedDict()  # (device, stream, R, P) -> (log_params, val, grad, info) staging buffers


def mll_grad_batched(kind: str, X, y, log_params, has_kv: bool, fixed_kv: float, noise: float,
                     max_batch: Optional[int] = None, reuse_buffers: bool = False
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """log p(y | theta_r) and d/d theta for R restarts -- BOBE/gp.py:385-398 via BOBE/optim.py:309.

    ``reuse_buffers=True``: the parameters are copied into, and the results returned in, buffers that persist per
    (device, stream, R, P).  The native call then sees the same pointers every time and replays its captured CUDA graph
    (one graph launch instead of several hundred kernel launches on the host).  The returned tensors are only valid until
    the next call with the same shape on the same stream -- for callers that fetch the results at once (the optimisers)."""
    X, y = _chk(X, "X"), _chk(y, "y").reshape(-1)
    n, d = X.shape
    dev = X.device
    if reuse_buffers:
        lp_in = log_params if isinstance(log_params, torch.Tensor) else torch.as_tensor(log_params, dtype=torch.float64)
        if lp_in.dim() == 1:
            lp_in = lp_in[None, :]
        R, P = lp_in.shape
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (di, torch.cuda.current_stream(di).cuda_stream, R, P)
        bufs = _mll_io_cache.get(key)
        if bufs is None:
            bufs = (torch.empty((R, P), dtype=torch.float64, device=dev), torch.empty(R, dtype=torch.float64, device=dev),
                    torch.empty((R, P), dtype=torch.float64, device=dev), torch.empty(R, dtype=torch.int32, device=dev))
            _mll_io_cache[key] = bufs
            while len(_mll_io_cache) > 8:
                _mll_io_cache.popitem(last=False)
        _mll_io_cache.move_to_end(key)
        lp, val, grad, info = bufs
        lp.copy_(lp_in, non_blocking=True)
    else:
        lp = _chk(log_params, "log_params")
        if lp.dim() == 1:
            lp = lp[None, :]
        R, P = lp.shape
        val = torch.empty(R, dtype=torch.float64, device=dev)
        grad = torch.empty((R, P), dtype=torch.float64, device=dev)
        info = torch.empty(R, dtype=torch.int32, device=dev)
    if max_batch is None:  # keep the workspace under ~1/3 of device memory
        per = lib.bobe_mll_grad_workspace_bytes(n, d, 1)
        free = torch.cuda.get_device_properties(dev).total_memory // 3
        max_batch = max(1, int(free // max(per, 1)))
    with torch.cuda.device(dev):
        for r0 in range(0, R, max_batch):
            r1 = min(R, r0 + max_batch)
            nbytes = lib.bobe_mll_grad_workspace_bytes(n, d, r1 - r0)
            ws = _workspace(nbytes, dev)
            check(lib.bobe_mll_grad_batched(_stream(), KIND[kind], X.data_ptr(), y.data_ptr(), n, d,
                                            lp[r0:r1].data_ptr(), r1 - r0, P, int(bool(has_kv)), float(fixed_kv),
                                            float(noise), val[r0:r1].data_ptr(), grad[r0:r1].data_ptr(),
                                            info[r0:r1].data_ptr(), ws.data_ptr(), ws.numel()),
                  "bobe_mll_grad_batched")
    return val, grad, info


def predict(kind: str, X, ls, kv: float, noise: float, Linv, alpha, Xq, y_mean: float, y_std: float,
            want_mean: bool = True, want_var: bool = True, standardised: bool = False):
    """Posterior mean / variance at Xq (M, d) -- BOBE/gp.py:450-493."""
    X, ls, Xq = _chk(X, "X"), _chk(ls, "ls"), _chk(Xq, "Xq")
    n, d = X.shape
    M = Xq.shape[0]
    dev = X.device
    mode = (_lib.PREDICT_MEAN if want_mean else 0) | (_lib.PREDICT_VAR if want_var else 0) | \
           (_lib.PREDICT_STANDARDISED if standardised else 0)
    mean = torch.empty(M, dtype=torch.float64, device=dev) if want_mean else None
    var = torch.empty(M, dtype=torch.float64, device=dev) if want_var else None
    with torch.cuda.device(dev):
        nbytes = lib.bobe_predict_workspace_bytes(n, d, M, mode)
        ws = _workspace(nbytes, dev)
        check(lib.bobe_predict(_stream(), KIND[kind], X.data_ptr(), n, d, ls.data_ptr(), float(kv), float(noise),
                               _chk(Linv, "Linv").data_ptr() if want_var else None,
                               _chk(alpha, "alpha").data_ptr() if want_mean else None, Xq.data_ptr(), M, float(y_mean),
                               float(y_std), mode, mean.data_ptr() if want_mean else None,
                               var.data_ptr() if want_var else None, ws.data_ptr(), ws.numel()), "bobe_predict")
    return mean, var


def linv_transpose(Linv, n: int) -> torch.Tensor:
    """Linv^T (npad, npad) from the padded inverse factor of ``factorize``."""
    Linv = _chk(Linv, "Linv")
    out = torch.empty_like(Linv)
    with torch.cuda.device(Linv.device):
        check(lib.bobe_linv_transpose(_stream(), Linv.data_ptr(), n, out.data_ptr()), "bobe_linv_transpose")
    return out


def predict_grad(kind: str, X, ls, kv: float, noise: float, Linv, LinvT, alpha, Xq, y_mean: float, y_std: float,
                 want_mean: bool = True, want_var: bool = True, standardised: bool = False):
    """(mean, var, dmean/dx, dvar/dx) at Xq (M, d) -- value_and_grad of BOBE/gp.py:450-489 w.r.t. the query point."""
    X, ls, Xq = _chk(X, "X"), _chk(ls, "ls"), _chk(Xq, "Xq")
    n, d = X.shape
    M = Xq.shape[0]
    dev = X.device
    mode = (_lib.PREDICT_MEAN if want_mean else 0) | (_lib.PREDICT_VAR if want_var else 0) | \
           (_lib.PREDICT_STANDARDISED if standardised else 0)
    new = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
    mean, dmean = (new(M), new(M, d)) if want_mean else (None, None)
    var, dvar = (new(M), new(M, d)) if want_var else (None, None)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(dev):
        ws = _workspace(lib.bobe_predict_grad_workspace_bytes(n, d, M), dev)
        check(lib.bobe_predict_grad(_stream(), KIND[kind], X.data_ptr(), n, d, ls.data_ptr(), float(kv), float(noise),
                                    ptr(_chk(Linv, "Linv")) if want_var else None,
                                    ptr(_chk(LinvT, "LinvT")) if want_var else None,
                                    ptr(_chk(alpha, "alpha")) if want_mean else None, Xq.data_ptr(), M, float(y_mean),
                                    float(y_std), mode, ptr(mean), ptr(var), ptr(dmean), ptr(dvar), ws.data_ptr(),
                                    ws.numel()), "bobe_predict_grad")
    return mean, var, dmean, dvar


def fantasy_var(kind: str, X, ls, kv: float, noise: float, Linv, y_std: float, Xmc, Xcand=None, reduce: str = "none"):
    """Fantasy variance at the MC points for each candidate -- BOBE/gp.py:552-576, acquisition.py:438-465."""
    X, ls, Xmc, Linv = _chk(X, "X"), _chk(ls, "ls"), _chk(Xmc, "Xmc"), _chk(Linv, "Linv")
    n, d = X.shape
    n_mc = Xmc.shape[0]
    if Xcand is not None:
        Xcand = _chk(Xcand, "Xcand")
        C = Xcand.shape[0]
    else:
        C = n_mc
    red = {"none": _lib.REDUCE_NONE, "mean": _lib.REDUCE_MEAN, "mean_sqrt": _lib.REDUCE_MEAN_SQRT}[reduce]
    dev = X.device
    out = torch.empty((C, n_mc) if red == _lib.REDUCE_NONE else (C,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib.bobe_fantasy_var_workspace_bytes(n, d, n_mc, C if Xcand is not None else 0)
        ws = _workspace(nbytes, dev)
        check(lib.bobe_fantasy_var(_stream(), KIND[kind], X.data_ptr(), n, d, ls.data_ptr(), float(kv), float(noise),
                                   Linv.data_ptr(), float(y_std), Xmc.data_ptr(), n_mc,
                                   Xcand.data_ptr() if Xcand is not None else None, C, red, out.data_ptr(),
                                   ws.data_ptr(), ws.numel()), "bobe_fantasy_var")
    return out


def fantasy_var_grad(kind: str, X, ls, kv: float, noise: float, Linv, LinvT, y_std: float, Xmc, Xcand,
                     reduce: str = "mean"):
    """WIPV ("mean") / WIPStd ("mean_sqrt") at the candidates and their gradients w.r.t. the candidate point --
    value_and_grad of BOBE/acquisition.py:438-440,463-465 as taken at BOBE/acquisition.py:400-412.  -> ((C,), (C, d))."""
    X, ls, Xmc, Xcand = _chk(X, "X"), _chk(ls, "ls"), _chk(Xmc, "Xmc"), _chk(Xcand, "Xcand")
    Linv, LinvT = _chk(Linv, "Linv"), _chk(LinvT, "LinvT")
    n, d = X.shape
    n_mc, C = Xmc.shape[0], Xcand.shape[0]
    red = {"mean": _lib.REDUCE_MEAN, "mean_sqrt": _lib.REDUCE_MEAN_SQRT}[reduce]
    dev = X.device
    out = torch.empty(C, dtype=torch.float64, device=dev)
    dout = torch.empty((C, d), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.bobe_fantasy_var_grad_workspace_bytes(n, d, n_mc, C), dev)
        check(lib.bobe_fantasy_var_grad(_stream(), KIND[kind], X.data_ptr(), n, d, ls.data_ptr(), float(kv), float(noise),
                                        Linv.data_ptr(), LinvT.data_ptr(), float(y_std), Xmc.data_ptr(), n_mc,
                                        Xcand.data_ptr(), C, red, out.data_ptr(), dout.data_ptr(), ws.data_ptr(),
                                        ws.numel()), "bobe_fantasy_var_grad")
    return out, dout


def chol_append(L, k, k_self: float) -> torch.Tensor:
    """fast_update_cholesky -- BOBE/gp.py:181-197."""
    L, k = _chk(L, "L"), _chk(k, "k").reshape(-1)
    n = L.shape[0]
    out = torch.empty((n + 1, n + 1), dtype=torch.float64, device=L.device)
    with torch.cuda.device(L.device):
        check(lib.bobe_chol_append(_stream(), L.data_ptr(), n, L.shape[1] if n else 1, k.data_ptr(), float(k_self),
                                   out.data_ptr(), n + 1), "bobe_chol_append")
    return out


def acq_ei(which: str, mean, var, best_y: float, zeta: float) -> torch.Tensor:
    """Negated EI / LogEI from standardised (mean, var) -- BOBE/acquisition.py:226-253,318-330."""
    mean, var = _chk(mean, "mean").reshape(-1), _chk(var, "var").reshape(-1)
    out = torch.empty_like(mean)
    with torch.cuda.device(mean.device):
        check(lib.bobe_acq_ei(_stream(), _lib.ACQ_EI if which == "ei" else _lib.ACQ_LOGEI, mean.data_ptr(),
                              var.data_ptr(), mean.numel(), float(best_y), float(zeta), out.data_ptr()), "bobe_acq_ei")
    return out


def svm_mask(sv, dual_coef, intercept: float, gamma: float, Xq, mean=None, var=None, minus_inf: float = -1e5,
             var_fill: float = 1e-12, want_decision: bool = False):
    """RBF-SVM decision function at Xq and the ``jnp.where`` mask of BOBE/clf_gp.py:173-205 applied IN PLACE to
    ``mean`` / ``var`` (either may be None).  Returns the decision values if asked for."""
    sv, dual_coef, Xq = _chk(sv, "sv"), _chk(dual_coef, "dual_coef").reshape(-1), _chk(Xq, "Xq")
    n_sv, d = sv.shape
    M = Xq.shape[0]
    dev = Xq.device
    dec = torch.empty(M, dtype=torch.float64, device=dev) if want_decision else None
    ptr = lambda t: _chk(t, "mean/var").data_ptr() if t is not None else None
    with torch.cuda.device(dev):
        ws = _workspace(lib.bobe_svm_mask_workspace_bytes(d, M), dev)
        check(lib.bobe_svm_mask(_stream(), sv.data_ptr(), n_sv, d, dual_coef.data_ptr(), float(intercept), float(gamma),
                                Xq.data_ptr(), M, float(minus_inf), float(var_fill), ptr(mean), ptr(var),
                                dec.data_ptr() if dec is not None else None, ws.data_ptr(), ws.numel()), "bobe_svm_mask")
    return dec
