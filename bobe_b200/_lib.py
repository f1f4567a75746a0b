"""ctypes binding of the C-ABI in ``include/bobe_b200.h``.

There is no CPU fallback: if the shared library has not been built the import of this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbobe_b200.so")

KERNEL_RBF, KERNEL_MATERN52 = 0, 1
PREDICT_MEAN, PREDICT_VAR, PREDICT_STANDARDISED = 1, 2, 4
REDUCE_NONE, REDUCE_MEAN, REDUCE_MEAN_SQRT = 0, 1, 2
ACQ_EI, ACQ_LOGEI = 0, 1

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); kept in one table so that tests can check it against the header
SIGNATURES = {
    "bobe_last_error_string": (C.c_char_p, []),
    "bobe_abi_version": (_i32, []),
    "bobe_npad": (_i64, [_i64]),
    "bobe_kernel_matrix": (_i32, [_vp, _i32, _vp, _i64, _vp, _i64, _i64, _vp, _f64, _f64, _i32, _vp, _i64]),
    "bobe_factorize_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "bobe_factorize": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _vp, _vp, _f64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                              _vp, _i64]),
    "bobe_cholesky_workspace_bytes": (_i64, [_i64, _i64]),
    "bobe_cholesky_batched": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "bobe_dist_sq": (_i32, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64]),
    "bobe_factor_append_workspace_bytes": (_i64, [_i64, _i64]),
    "bobe_factor_append": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _i64, _vp, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _i64]),
    "bobe_mll_grad_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "bobe_mll_grad_batched": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _i32, _f64, _f64, _vp, _vp,
                                     _vp, _vp, _i64]),
    "bobe_predict_workspace_bytes": (_i64, [_i64, _i64, _i64, _i32]),
    "bobe_predict": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _f64, _f64, _vp, _vp, _vp, _i64, _f64, _f64, _i32, _vp,
                            _vp, _vp, _i64]),
    "bobe_linv_transpose": (_i32, [_vp, _vp, _i64, _vp]),
    "bobe_predict_grad_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "bobe_predict_grad": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _f64, _f64, _vp, _vp, _vp, _vp, _i64, _f64, _f64, _i32,
                                 _vp, _vp, _vp, _vp, _vp, _i64]),
    "bobe_fantasy_var_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64]),
    "bobe_fantasy_var": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _f64, _f64, _vp, _f64, _vp, _i64, _vp, _i64, _i32,
                                _vp, _vp, _i64]),
    "bobe_fantasy_var_grad_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64]),
    "bobe_fantasy_var_grad": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _f64, _f64, _vp, _vp, _f64, _vp, _i64, _vp, _i64,
                                     _i32, _vp, _vp, _vp, _i64]),
    "bobe_chol_append": (_i32, [_vp, _vp, _i64, _i64, _vp, _f64, _vp, _i64]),
    "bobe_acq_ei": (_i32, [_vp, _i32, _vp, _vp, _i64, _f64, _f64, _vp]),
    "bobe_svm_mask_workspace_bytes": (_i64, [_i64, _i64]),
    "bobe_svm_mask": (_i32, [_vp, _vp, _i64, _i64, _vp, _f64, _f64, _vp, _i64, _f64, _f64, _vp, _vp, _vp, _vp, _i64]),
    "bobe_bench_trmm_sumsq": (_i32, [_vp, _vp, _i64, _vp, _i64, _f64, _vp]),
}


class BobeNativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"bobe_b200: native library {LIB_PATH} is missing. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `python bobe_b200/build.py`). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str):
    if rc != 0:
        msg = lib.bobe_last_error_string()
        raise BobeNativeError(f"{what} failed with status {rc}: {msg.decode() if msg else '?'}")
