"""Ad-hoc GPU parity check of the raw ops against the oracle (development aid; the real tests are in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import gp_oracle as O
from bobe_b200 import ops

dev = "cuda"
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)

def relerr(a, b, scale=1.0):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), scale)))

def run(n, d, kernel, M, ell, R=3):
    X, y = O.synthetic_training_set(n, d)
    gp = O.OracleGP(X, y, kernel=kernel, lengthscales=np.full(d, ell))
    Xq = O.synthetic_queries(M, d)
    ls = T(gp.lengthscales)
    K = ops.kernel_matrix(gp.kernel_name, T(X), T(X), ls, gp.kernel_variance, gp.noise, True).cpu().numpy()
    Kref = gp.kernel(X, X, gp.lengthscales, gp.kernel_variance, gp.noise, True)
    print(f"[{kernel} n={n} d={d}] K relerr {relerr(K, Kref):.2e}")
    L, Linv, alpha, logdet, quad, info = ops.factorize(gp.kernel_name, T(X), T(gp.train_y), ls[None], T([gp.kernel_variance]), gp.noise)
    L = L[0, :n, :n].cpu().numpy(); al = alpha[0, :n].cpu().numpy()
    print(f"   L relerr {relerr(L, gp.cholesky, np.abs(gp.cholesky).max()):.2e} alpha rel {np.linalg.norm(al - gp.alphas.ravel())/np.linalg.norm(gp.alphas):.2e} info {info.item()}"
          f" logdet {logdet.item():.10f} ref {np.sum(np.log(np.diag(gp.cholesky))):.10f}")
    mean, var = ops.predict(gp.kernel_name, T(X), ls, gp.kernel_variance, gp.noise, Linv[0], alpha[0], T(Xq), gp.y_mean, gp.y_std)
    mref, vref = gp.predict_mean_batched(Xq), gp.predict_var_batched(Xq)
    print(f"   mean err {relerr(mean.cpu().numpy(), mref, gp.y_std):.2e}  var err {relerr(var.cpu().numpy(), vref, gp.y_std**2):.2e}")
    x0 = O.synthetic_restarts(gp, R)
    val, grad, info = ops.mll_grad_batched(gp.kernel_name, T(X), T(gp.train_y), T(x0), True, 1.0, gp.noise)
    for r in range(R):
        v, g = gp.neg_mll_and_grad(x0[r]); lp, lg = gp.log_prior_and_grad(x0[r])
        vref, gref = -v - lp, -g - lg
        gv = grad[r].cpu().numpy()
        print(f"   r{r}: mll {val[r].item():.9e} ref {vref:.9e} rel {abs(val[r].item()-vref)/max(abs(vref), n):.2e} grad err {relerr(gv, gref, np.abs(gref).max() if np.all(np.isfinite(gref)) else 1):.2e} info {info[r].item()}")
    nmc = 96
    mc = O.synthetic_queries(nmc, d, seed=5)
    fv = ops.fantasy_var(gp.kernel_name, T(X), ls, gp.kernel_variance, gp.noise, Linv[0], gp.y_std, T(mc), None, "none").cpu().numpy()
    ktm = gp.kernel(X, mc, gp.lengthscales, gp.kernel_variance, gp.noise, False)
    ref = np.stack([gp.fantasy_var(mc[c], mc, ktm) for c in range(8)])
    print(f"   fantasy(self) err {relerr(fv[:8], ref, gp.y_std**2):.2e}")
    cand = O.synthetic_queries(5, d, seed=6)
    fv2 = ops.fantasy_var(gp.kernel_name, T(X), ls, gp.kernel_variance, gp.noise, Linv[0], gp.y_std, T(mc), T(cand), "mean").cpu().numpy()
    ref2 = np.array([np.mean(gp.fantasy_var(cand[c], mc, ktm)) for c in range(5)])
    print(f"   wipv err {relerr(fv2, ref2, gp.y_std**2):.2e}")
    k = gp._k12(cand[0]).ravel()
    La = ops.chol_append(T(gp.cholesky), T(k), gp.kernel_variance + gp.noise).cpu().numpy()
    print(f"   chol_append err {relerr(La, O.fast_update_cholesky(gp.cholesky, k, gp.kernel_variance + gp.noise), 1.0):.2e}")
    ms, vs = gp.predict_batched(Xq[:256])
    for w in ("ei", "logei"):
        o = ops.acq_ei(w, T(ms), T(vs.ravel()), float(gp.train_y.max()), 0.01).cpu().numpy()
        r = (O.ei_values if w == "ei" else O.logei_values)(ms, vs, float(gp.train_y.max()), 0.01)
        print(f"   {w} err {relerr(o, r, 1e-300):.2e}")

if __name__ == "__main__":
    run(100, 2, "rbf", 1000, 0.3)
    run(500, 4, "rbf", 3000, 0.5)
    run(300, 3, "matern", 777, 0.7)
    run(1000, 8, "matern", 5000, 1.0, R=4)
    torch.cuda.synchronize()
    print("done")
