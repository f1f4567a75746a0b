"""End-to-end GP.fit timing (config C): 64 restarts, n=2000, d=16 Matern, lock-step L-BFGS-B through the batched call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP, optim
from oracle import gp_oracle as O
n, d, R = int(sys.argv[1]) if len(sys.argv) > 1 else 2000, 16, int(sys.argv[2]) if len(sys.argv) > 2 else 64
X, y = O.synthetic_training_set(n, d)
for opt in ("scipy", "optax"):
    gp = GP(X, y, kernel="matern", lengthscales=np.ones(d), optimizer=opt)
    ref = O.OracleGP(X, y, kernel="matern", lengthscales=np.ones(d))
    x0 = O.synthetic_restarts(ref, R)
    calls = {"n": 0, "pts": 0}
    orig = gp.neg_mll_and_grad_batched
    def counted(lp, _o=orig):
        calls["n"] += 1; calls["pts"] += len(lp)
        return _o(lp)
    gp.neg_mll_and_grad_batched = counted
    gp.fit(x0[:4], maxiter=2)  # warm-up
    calls["n"] = calls["pts"] = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = gp.fit(x0, maxiter=15)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{opt}: fit R={R} maxiter=15: {dt*1e3:.0f} ms, {calls['n']} batched calls, {calls['pts']} evals -> {calls['pts']/dt:.0f} evals/s end to end, "
          f"{dt*1e3/max(calls['n'],1):.1f} ms per call; best mll {res['mll']:.4f}")
