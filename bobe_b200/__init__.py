"""bobe_b200 -- B200-native (sm_100a) GP surrogate hot path, drop-in for BOBE's ``gp.py`` / ``acquisition.py``.

Importing the package loads the CUDA shared library; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError loudly if the native library is missing)
from . import ops  # noqa: F401

__version__ = "0.1.0"
