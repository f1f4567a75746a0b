"""Per-shard timing of the bench's log-ML+grad round at 8 restarts per rank (what each rank of the N = 8 run executes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP
from oracle import gp_oracle as O
X, y = O.synthetic_training_set(2000, 16)
gp = GP(X, y, kernel="matern", lengthscales=np.ones(16))
ref = O.OracleGP(X, y, kernel="matern", lengthscales=np.ones(16))
x0 = O.synthetic_restarts(ref, 64)
gp._ensure_factor()
out = []
for r in range(8):
    lp = x0[8 * r:8 * r + 8]
    for _ in range(4): gp.neg_mll_and_grad_batched(lp)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): gp.neg_mll_and_grad_batched(lp)
    out.append((time.perf_counter() - t0) / 10 * 1e3)
print({k: v for k, v in os.environ.items() if k.startswith("BOBE_")}, " ".join(f"{t:.2f}" for t in out), "max", f"{max(out):.2f}")
