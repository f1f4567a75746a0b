"""Sanity check beyond the reference's sizes: n = 8000 / 12000 training points against torch float64 (cuSOLVER / cuBLAS)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP
from oracle import gp_oracle as O
for n, d in ((8000, 8), (12000, 10)):
    X, y = O.synthetic_training_set(n, d)
    ls = np.full(d, 0.6)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    gp = GP(X, y, kernel="matern", lengthscales=ls, noise=1e-6)
    gp._ensure_factor(); torch.cuda.synchronize(); t_fac = time.perf_counter() - t0
    Xq = O.synthetic_queries(4096, d)
    m, v = gp.predict_mean_batched(Xq), gp.predict_var_batched(Xq)
    # torch float64 reference
    Xt, Xqt = torch.as_tensor(X, device="cuda"), torch.as_tensor(Xq, device="cuda")
    lt = torch.as_tensor(ls, device="cuda")
    def kern(a, b):
        r = torch.cdist(a / lt, b / lt).clamp_min(1e-15)
        return (1 + 5 ** 0.5 * r + 5.0 / 3.0 * r * r) * torch.exp(-5 ** 0.5 * r)
    K = kern(Xt, Xt); K.diagonal().fill_(1.0 + 1e-6)
    ys = torch.as_tensor((y - y.mean()) / y.std(), device="cuda")
    L = torch.linalg.cholesky(K)
    alpha = torch.cholesky_solve(ys, L)
    ks = kern(Xt, Xqt)
    mt = (ks.T @ alpha).ravel() * y.std() + y.mean()
    vv = torch.linalg.solve_triangular(L, ks, upper=False)
    vt = ((1.0 + 1e-6) - (vv * vv).sum(0)).clamp_min(1e-12) * y.std() ** 2
    em = float((torch.as_tensor(m, device="cuda") - mt).abs().max()) / y.std()
    ev = float((torch.as_tensor(v, device="cuda") - vt).abs().max()) / y.std() ** 2
    print(f"n={n} d={d}: factorise {t_fac*1e3:.1f} ms (incl. host), mean err {em:.2e} (y_std units), var err {ev:.2e}; "
          f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    del K, L, vv, ks, gp
    torch.cuda.empty_cache()
