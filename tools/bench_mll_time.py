"""Time the log-ML+grad round on the BENCH workload (the 64 synthetic restarts of bench.py, some ill-conditioned) at the
ABI and through GP.neg_mll_and_grad_batched (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP, ops
from oracle import gp_oracle as O
from tools.factor_ab import ev_time
X, y = O.synthetic_training_set(2000, 16)
gp = GP(X, y, kernel="matern", lengthscales=np.ones(16))
ref = O.OracleGP(X, y, kernel="matern", lengthscales=np.ones(16))
x0 = O.synthetic_restarts(ref, 64)
gp._ensure_factor()
res = []
for R in (64, 8):
    lp = torch.as_tensor(x0[:R], device="cuda")
    t = ev_time(lambda: ops.mll_grad_batched("matern", gp._X_dev, gp._y_dev, lp, True, 1.0, 1e-8), iters=6, warm=2)
    for _ in range(3): gp.neg_mll_and_grad_batched(x0[:R])
    t0 = time.perf_counter()
    for _ in range(8): gp.neg_mll_and_grad_batched(x0[:R])
    tw = (time.perf_counter() - t0) / 8 * 1e3
    res.append(f"R={R}: ABI back-to-back {t:.3f} ms, GP.neg_mll_and_grad_batched wall {tw:.3f} ms")
# per-row cost: which restarts carry the gated correction work
lp1 = torch.as_tensor(x0[:64], device="cuda")
print({k: v for k, v in os.environ.items() if k.startswith("BOBE_")}, " | ".join(res))
