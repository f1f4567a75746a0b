import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _load_build_module(root):
    """bobe_b200/build.py loaded by PATH: importing it as bobe_b200.build would run bobe_b200/__init__, which refuses
    to import without the native library -- the very thing this module creates."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bobe_b200_build", os.path.join(root, "bobe_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the native library must exist before bobe_b200 is imported (no CPU fallback); nvcc cross-compiles here
    _load_build_module(ROOT).build_native(force=False)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def mixed_err(a, ref, scale):
    """max |a - ref| / max(|ref|, scale)  -- the mixed tolerance of SURVEY.md 8c."""
    import numpy as np
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), scale))) if a.size else 0.0
