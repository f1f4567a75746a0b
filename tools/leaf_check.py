"""BOBE_LEAF_PANEL4 = 0 / 1 must give bitwise equal results (same operations per element in the same order)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
os.chdir(ROOT)
from test_gpu_parity import _run_with_env, _KERNEL_VARIANT_SNIPPET
a = _run_with_env({"BOBE_LEAF_PANEL4": "0"}, _KERNEL_VARIANT_SNIPPET)
b = _run_with_env({"BOBE_LEAF_PANEL4": "1"}, _KERNEL_VARIANT_SNIPPET)
print("bitwise equal:", a == b, a[:16], b[:16])
