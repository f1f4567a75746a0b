import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.optimize import minimize
from bobe_b200 import GP
from oracle import gp_oracle as O
rng = np.random.RandomState(42); X = rng.uniform(0, 1, size=(30, 2)); y = -np.sum((X - 0.5) ** 2, axis=1).reshape(-1, 1)
gp = GP(train_x=X, train_y=y, noise=1e-6, kernel="matern", lengthscale_prior="DSLP")
ref = O.OracleGP(X, y, noise=1e-6, kernel="matern", lengthscale_prior="DSLP")
x0 = np.log(gp.get_hyperparams())
cnt = [0]
def vg(x):
    v, g = gp.neg_mll_and_grad(x); vr, gr = ref.neg_mll_and_grad(x); cnt[0] += 1
    print(cnt[0], x, v, vr, abs(v - vr), np.abs(g - gr).max())
    return v, g
res = minimize(vg, x0, method="L-BFGS-B", jac=True, bounds=[tuple(b) for b in gp.hyperparam_bounds.T], options={"maxiter": 200})
print(res.message, res.success, res.fun, res.x)
