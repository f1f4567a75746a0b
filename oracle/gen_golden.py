"""Generate the golden vectors under tests/golden/ from the NumPy oracle (and the mpmath truth for the small
case).  TEST INFRASTRUCTURE.  Run from the repo root:  python -m oracle.gen_golden

JAX is not installable in this image, so these are outputs of the restatement at the full BASELINE shapes; the restatement
itself is held to the reference's own source (executed under a NumPy stand-in: oracle/gen_reference_vectors.py ->
tests/golden/reference_source_vectors.npz) and to the mathematics in tests/test_oracle.py.
Inputs are regenerated from seeds (oracle.gp_oracle.synthetic_*), so only outputs are stored.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gp_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name: (n, d, kernel, lengthscale, M queries, R restarts, n_mc, C candidates)
CASES = {
    "A_banana_rbf_n100_d2": (100, 2, "rbf", 0.3, 256, 4, 64, 8),
    "B_rbf_n500_d2": (500, 2, "rbf", 0.3, 256, 4, 64, 8),  # worst-conditioned BASELINE shape: cond(K) = 1.8e10
    "B_rbf_n500_d4": (500, 4, "rbf", 0.5, 256, 4, 64, 8),
    "B_rbf_n500_d6": (500, 6, "rbf", 0.5, 256, 3, 64, 8),
    "M_matern_n300_d3": (300, 3, "matern", 0.7, 256, 4, 64, 8),
    "D_rbf_n1500_d27": (1500, 27, "rbf", 2.0, 128, 2, 32, 4),
    "H_matern_n2000_d16": (2000, 16, "matern", 1.0, 128, 3, 32, 4),
}


def make_case(name):
    n, d, kern, ell, M, R, n_mc, C = CASES[name]
    X, y = O.synthetic_training_set(n, d)
    gp = O.OracleGP(X, y, kernel=kern, lengthscales=np.full(d, ell))
    Xq = O.synthetic_queries(M, d)
    x0 = O.synthetic_restarts(gp, R)
    mc = O.synthetic_queries(n_mc, d, seed=5)
    cand = O.synthetic_queries(C, d, seed=6)
    return gp, X, y, Xq, x0, mc, cand


def main(names=None):
    """Regenerate all cases, or only the named ones:  python -m oracle.gen_golden B_rbf_n500_d2"""
    os.makedirs(OUT, exist_ok=True)
    for name in (names or CASES):
        gp, X, y, Xq, x0, mc, cand = make_case(name)
        out = {}
        out["mean"] = gp.predict_mean_batched(Xq)
        out["var"] = gp.predict_var_batched(Xq)
        ms, vs = gp.predict_batched(Xq)
        out["mean_std"], out["var_std"] = ms, vs.ravel()
        vals, grads = [], []
        for r in range(x0.shape[0]):
            v, g = gp.neg_mll_and_grad(x0[r])
            vals.append(v)
            grads.append(g)
        out["neg_mll"], out["neg_mll_grad"] = np.array(vals), np.array(grads)
        out["fantasy"] = gp.fantasy_var_shared(cand, mc)
        out["wipv_self"] = O.wipv_values(gp, mc, mc)
        out["wipstd_self"] = O.wipv_values(gp, mc, mc, std=True)
        best = float(gp.train_y.max())
        out["ei"] = O.ei_values(ms, vs, best, 0.01)
        out["logei"] = O.logei_values(ms, vs, best, 0.01)
        out["logdet"] = np.array(np.sum(np.log(np.diag(gp.cholesky))))
        out["alpha"] = gp.alphas.ravel()
        if X.shape[0] <= 500:  # exact-arithmetic values, so that parity can be read against the oracle's own noise
            from oracle.truth_mp import TruthGP
            T = TruthGP(gp.kernel_name, X, gp.train_y, gp.lengthscales, gp.kernel_variance, gp.noise)
            mt, vt = T.predict(Xq[:32])
            out["truth_mean_std"], out["truth_var_raw"] = mt, vt
            out["truth_mll"] = np.array(T.mll())  # at the current hyper-parameters = restart 0
            if X.shape[0] <= 100:
                out["truth_mll_grad"] = T.mll_grad()
            # ... and at every restart row (random hyper-parameters: often far worse conditioned than row 0), so that
            # the CUDA log-ML of every row can be read against the exact value and the oracle's own rounding noise
            rows = []
            for r in range(x0.shape[0]):
                ls_r, kv_r, _ = gp._parse_hyperparams(x0[r])
                rows.append(float(TruthGP(gp.kernel_name, X, gp.train_y, ls_r, kv_r, gp.noise).mll()))
            out["truth_mll_rows"] = np.array(rows)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1:] or None)
