"""NCCL check of bobe_b200.dist on real GPUs (run under torchrun, one rank per GPU): sharded predict / log-ML+grad /
fit / acquisition must equal the single-rank results."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as tdist
rank, lrank, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lrank)
tdist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
from bobe_b200 import GP, dist
from oracle import gp_oracle as O
n, d = 400, 4
X, y = O.synthetic_training_set(n, d)
gp = GP(X, y, kernel="matern", lengthscales=np.full(d, 0.7))
ref = O.OracleGP(X, y, kernel="matern", lengthscales=np.full(d, 0.7))
Xq = O.synthetic_queries(10_001, d)
m, v = dist.predict_sharded(gp, Xq)
m1, v1 = gp.predict_mean_var_batched(Xq)
# the mean is bitwise batch-invariant; the variance k** - |v|^2 may differ in the last bit of |v|^2 between the fused and
# the row-split reductions (shards of different size take different paths): compare on the scale of k** y_std^2
assert np.array_equal(m, m1), ("predict_sharded mean", float(np.abs(m - m1).max()))
assert float(np.abs(v - v1).max()) <= 1e-13 * ref.y_std ** 2, ("predict_sharded var", float(np.abs(v - v1).max()))
x0 = O.synthetic_restarts(ref, 7)
val, grad = dist.mll_grad_sharded(gp, x0)
val1, grad1 = gp.neg_mll_and_grad_batched(x0)
assert np.allclose(val, val1, rtol=1e-12, equal_nan=True) and np.allclose(grad, grad1, rtol=1e-10, atol=1e-12, equal_nan=True), "mll_grad_sharded"
res = dist.fit_sharded(lambda chunk: gp.fit(chunk, maxiter=10), x0)
res1 = gp.fit(x0, maxiter=10)
assert np.isfinite(res['mll']) and res['mll'] >= res1['mll'] - 1e-6 * abs(res1['mll']), ("fit_sharded", res['mll'], res1['mll'])
mc = O.synthetic_queries(96, d, seed=5)
cand = O.synthetic_queries(13, d, seed=6)
a = dist.acquisition_sharded(lambda c: gp.fantasy_acquisition(mc, c), cand)
a1 = gp.fantasy_acquisition(mc, cand)
assert np.allclose(a, a1, rtol=1e-12), "acquisition_sharded"
# WIPV / WIPStd with the MC columns sharded (dist.wipv_sharded): equal to the single-GPU value up to the regrouping of the mean
mc2 = O.synthetic_queries(1001, d, seed=7)
for std in (False, True):
    w = dist.wipv_sharded(gp, mc2, cand, std=std)
    w1 = gp.fantasy_acquisition(mc2, cand, std=std)
    assert w.shape == (13,) and np.allclose(w, w1, rtol=1e-11, atol=0), ("wipv_sharded", float(np.abs(w / w1 - 1).max()))
ws_self = dist.wipv_sharded(gp, mc, None)
assert np.allclose(ws_self, gp.fantasy_acquisition(mc, None), rtol=1e-11), "wipv_sharded self"
gathered = [None] * ws
tdist.all_gather_object(gathered, w.tobytes())
assert all(g == gathered[0] for g in gathered), "wipv_sharded must be bit-identical on every rank"
tdist.barrier()
if rank == 0:
    print(f"dist_check ok on {ws} ranks: predict {m.shape}, mll {val.shape}, fit mll {res['mll']:.6f} (single-rank {res1['mll']:.6f}), acq {a.shape}, wipv_sharded max rel dev {float(np.abs(w / w1 - 1).max()):.1e}")
tdist.destroy_process_group()
