"""Short single-GPU workload for ncu captures: a few predict chunks and one batched log-ML+grad round."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP, ops
from oracle import gp_oracle as O
what = sys.argv[1] if len(sys.argv) > 1 else "predict"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
X, y = O.synthetic_training_set(2000, 16)
gp = GP(X, y, kernel="matern", lengthscales=np.ones(16))
if what == "predict":
    Xq = torch.rand(3 * 18944, 16, dtype=torch.float64, device="cuda")
    for _ in range(2):
        m, v = gp.predict_mean_var_batched(Xq)
else:
    ref = O.OracleGP(X, y, kernel="matern", lengthscales=np.ones(16))
    lp = torch.as_tensor(O.synthetic_restarts(ref, R), device="cuda")
    for _ in range(2):
        val, grad, info = ops.mll_grad_batched("matern", gp._X_dev, gp._y_dev, lp, True, 1.0, 1e-8)
torch.cuda.synchronize()
print("done", what)
