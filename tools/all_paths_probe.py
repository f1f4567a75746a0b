"""Small end-to-end pass over every kernel family checked against the oracle (compute-sanitizer is closed on this pool): sizes chosen so that ragged
tiles, the look-ahead scheme (batch <= 8), the single-stream scheme (batch > 8), the row-split and shared-panel variance
launches and the fantasy / gradient paths all run, in seconds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP, ops, EI, WIPV
from oracle import gp_oracle as O

for n, d, R in ((300, 3, 2), (517, 4, 10)):
    X, y = O.synthetic_training_set(n, d)
    gp = GP(X, y, kernel="matern", lengthscales=np.full(d, 0.7))
    ref = O.OracleGP(X, y, kernel="matern", lengthscales=np.full(d, 0.7))
    lp = O.synthetic_restarts(ref, R)
    v, g = gp.neg_mll_and_grad_batched(lp)
    vr = np.array([ref.neg_mll_and_grad(r)[0] for r in lp[:2]])
    assert np.allclose(v[:2], vr, rtol=1e-9), (v[:2], vr)
    Xq = O.synthetic_queries(148 * 128 + 5000, d)
    m, var = gp.predict_mean_var_batched(Xq)
    idx = np.arange(0, Xq.shape[0], 997)
    assert np.allclose(var[idx], ref.predict_var_batched(Xq[idx]), rtol=1e-6, atol=1e-9 * ref.y_std ** 2)
    mc = O.synthetic_queries(700, d, seed=3)
    w = gp.fantasy_acquisition(mc, O.synthetic_queries(9, d, seed=4))
    val, grad = WIPV().value_and_grad_batched(O.synthetic_queries(3, d, seed=5), gp, mc_points=mc)
    e, ge = EI().value_and_grad_batched(O.synthetic_queries(5, d, seed=6), gp, float(ref.train_y.max()), 0.01)
    gp.update(O.synthetic_queries(2, d, seed=8), np.array([[0.1], [0.2]]))
    gp.predict_mean_batched(Xq[:100])
    torch.cuda.synchronize()
    print("ok", n, d, R)
