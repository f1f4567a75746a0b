"""Optional ``jax.ffi`` registration of the C-ABI.

UNTESTED IN THIS ENVIRONMENT: JAX is not installed here or on the GPU boxes and the XLA FFI headers are absent,
so ``csrc/xla_ffi_shim.cc`` cannot be compiled.  The module imports cleanly without JAX; ``register()`` raises a
clear error instead of silently doing nothing.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIM_PATH = os.path.join(_HERE, "lib", "libbobe_xla_ffi.so")
TARGETS = {"bobe_kernel_matrix": "BobeKernelMatrix", "bobe_factorize": "BobeFactorize",
           "bobe_mll_grad": "BobeMllGrad", "bobe_predict": "BobePredict", "bobe_fantasy_var": "BobeFantasyVar",
           "bobe_fantasy_var_grad": "BobeFantasyVarGrad",
           "bobe_predict_grad": "BobePredictGrad", "bobe_linv_transpose": "BobeLinvTranspose",
           "bobe_factor_append": "BobeFactorAppend", "bobe_svm_mask": "BobeSvmMask"}


def available() -> bool:
    try:
        import jax  # noqa: F401
    except Exception:
        return False
    return os.path.exists(SHIM_PATH)


def register():
    """Register every handler of the shim as a CUDA FFI target."""
    try:
        import jax
    except Exception as e:  # pragma: no cover - JAX absent in this image
        raise ImportError("bobe_b200.jax_ffi needs JAX (not installed in this environment)") from e
    if not os.path.exists(SHIM_PATH):
        raise ImportError(f"{SHIM_PATH} not built; see the build line at the top of csrc/xla_ffi_shim.cc")
    shim = ctypes.CDLL(SHIM_PATH)
    for target, symbol in TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(shim, symbol)), platform="CUDA")


def predict(X, ls, Linv, alpha, Xq, *, kind, kv, noise, y_mean, y_std, mode):  # pragma: no cover
    """jit-compatible posterior mean/variance: BOBE/gp.py:450-493 as one custom call."""
    import jax
    import jax.numpy as jnp
    import numpy as np
    from ._lib import lib
    n, d = X.shape
    M = Xq.shape[0]
    ws = int(lib.bobe_predict_workspace_bytes(n, d, M, mode))
    out = (jax.ShapeDtypeStruct((M,), jnp.float64), jax.ShapeDtypeStruct((M,), jnp.float64),
           jax.ShapeDtypeStruct((ws,), jnp.uint8))
    mean, var, _ = jax.ffi.ffi_call("bobe_predict", out)(
        X, ls, Linv, alpha, Xq, kind=np.int64(kind), kv=np.float64(kv), noise=np.float64(noise),
        y_mean=np.float64(y_mean), y_std=np.float64(y_std), mode=np.int64(mode))
    return mean, var


def neg_mll_value_and_grad(X, y, log_params, *, kind, has_kv, fixed_kv, noise):  # pragma: no cover
    """Data term of GP.neg_mll with its gradient from one launch sequence; wrap in jax.custom_vjp so that
    jax.value_and_grad (BOBE/optim.py:118,211,309) keeps working."""
    import jax
    import jax.numpy as jnp
    import numpy as np
    from ._lib import lib
    n, d = X.shape
    R, P = log_params.shape
    ws = int(lib.bobe_mll_grad_workspace_bytes(n, d, R))
    out = (jax.ShapeDtypeStruct((R,), jnp.float64), jax.ShapeDtypeStruct((R, P), jnp.float64),
           jax.ShapeDtypeStruct((R,), jnp.int32), jax.ShapeDtypeStruct((ws,), jnp.uint8))
    val, grad, info, _ = jax.ffi.ffi_call("bobe_mll_grad", out)(
        X, y, log_params, kind=np.int64(kind), has_kv=np.int64(has_kv), fixed_kv=np.float64(fixed_kv),
        noise=np.float64(noise))
    return -val, -grad, info
