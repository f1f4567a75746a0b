"""Sub-batch policy sweep: log-ML+grad round for several R with the stream count forced by the environment (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bobe_b200 import ops
from tools.factor_ab import ev_time, data
import numpy as np
from oracle import gp_oracle as O
Xh, yh = O.synthetic_training_set(2000, 16)
ref = O.OracleGP(Xh, yh, kernel="matern", lengthscales=np.ones(16))
x0 = O.synthetic_restarts(ref, 64)  # the bench workload: rows drawn as BOBE/pool.py:277-284 does, some ill-conditioned
X = torch.as_tensor(ref.train_x, device="cuda"); y = torch.as_tensor(ref.train_y.ravel(), device="cuda")
out = []
for R in (8, 16, 32, 64):
    lp = torch.as_tensor(x0[:R], device="cuda")
    t = ev_time(lambda: ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8), iters=5, warm=2)
    out.append(f"R={R}: {t:.2f}")
print(os.environ.get("BOBE_MLL_STREAMS"), os.environ.get("BOBE_LOOKAHEAD_MAX", "8"), " ".join(out))
