"""Accuracy probe: backward errors of the recursive Cholesky/inverse vs LAPACK at ill-conditioned shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.linalg as sla
from oracle import gp_oracle as O
from bobe_b200 import ops
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
for (n, d, kern, ell) in [(100, 2, "rbf", 0.3), (500, 2, "rbf", 0.3), (500, 6, "rbf", 0.5)]:
    X, y = O.synthetic_training_set(n, d)
    gp = O.OracleGP(X, y, kernel=kern, lengthscales=np.full(d, ell))
    K = gp.kernel(X, X, gp.lengthscales, gp.kernel_variance, gp.noise, True)
    Kl = K.astype(np.longdouble)
    L, Linv, alpha, logdet, quad, info = ops.factorize(gp.kernel_name, T(X), T(gp.train_y), T(gp.lengthscales)[None], T([1.0]), gp.noise)
    Lg = L[0, :n, :n].cpu().numpy(); Li = Linv[0, :n, :n].cpu().numpy(); al = alpha[0, :n].cpu().numpy()
    Ll = gp.cholesky
    be = lambda L_: float(np.abs(Kl - L_.astype(np.longdouble) @ L_.astype(np.longdouble).T).max() / np.abs(K).max())
    print(f"n={n} d={d} {kern}: cond {np.linalg.cond(K):.2e}")
    print(f"  backward err K-LL^T: mine {be(Lg):.2e} lapack {be(Ll):.2e}")
    Lil = sla.solve_triangular(Ll, np.eye(n), lower=True)
    ri = lambda Li_, L_: float(np.abs(np.eye(n) - Li_.astype(np.longdouble) @ L_.astype(np.longdouble)).max())
    print(f"  |I - Linv L|: mine {ri(Li, Lg):.2e} lapack(trsm) {ri(Lil, Ll):.2e};  |I - L Linv| mine {ri(Lg, Li):.2e} lapack {ri(Ll, Lil):.2e}")
    res = lambda a_: float(np.abs(gp.train_y.ravel().astype(np.longdouble) - Kl @ a_.astype(np.longdouble)).max())
    print(f"  residual |y-K alpha|: mine {res(al):.2e} lapack {res(gp.alphas.ravel()):.2e}; alpha scale {np.abs(al).max():.2e}")
    Kinv_m = Li.T @ Li; Kinv_l = Lil.T @ Lil
    rk = lambda Ki_: float(np.abs(np.eye(n) - Kl @ Ki_.astype(np.longdouble)).max())
    print(f"  |I - K Kinv|: mine {rk(Kinv_m):.2e} lapack {rk(Kinv_l):.2e}")
