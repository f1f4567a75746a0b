"""bobe_b200 -- B200-native (sm_100a) GP surrogate hot path, drop-in for BOBE's ``gp.py`` / ``acquisition.py``.

Importing the package loads the CUDA shared library; there is no CPU fallback.
Public names follow the reference's ``BOBE/__init__.py:70-91`` for the path this package covers.
"""
from . import _lib  # noqa: F401  (raises ImportError loudly if the native library is missing)
from . import ops  # noqa: F401
from .gp import GP, rbf_kernel, matern_kernel, kernel_diag, fast_update_cholesky, dist_sq, gp_mll  # noqa: F401
from .acquisition import (AcquisitionFunction, EI, LogEI, WIPV, WIPStd, get_mc_samples, get_mc_points,  # noqa: F401
                          ACQUISITIONS)
from .optim import optimize_scipy, optimize_optax, optimize_optax_vmap, scale_to_unit, scale_from_unit  # noqa: F401
from .batching import SurrogatePool, lax_map  # noqa: F401
from .clf_gp import GPwithClassifier  # noqa: F401

__version__ = "0.1.0"
