"""Two kernel-matrix launches for ncu: the lower-triangle build of a 64-restart log-ML round and the K* panel of config B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP, ops
from oracle import gp_oracle as O
what = sys.argv[1]
if what == "mll":
    n, d, R = 2000, 16, 64
    X = torch.rand(n, d, dtype=torch.float64, device="cuda")
    y = torch.randn(n, dtype=torch.float64, device="cuda")
    lp = torch.log(torch.cat([0.5 + torch.rand(R, d, dtype=torch.float64, device="cuda"), torch.ones(R, 1, dtype=torch.float64, device="cuda")], 1))
    for _ in range(2):
        ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8)
else:
    X, y = O.synthetic_training_set(500, 4)
    gp = GP(X, y, kernel="rbf", lengthscales=np.full(4, 0.3), device="cuda")
    Xq = torch.as_tensor(O.synthetic_queries(100_000, 4), device="cuda")
    for _ in range(2):
        gp.predict_mean_var_batched(Xq)
torch.cuda.synchronize()
