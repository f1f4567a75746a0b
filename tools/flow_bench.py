"""Wall-clock of the host-level flows one BO iteration is made of (development aid; small-n regime of the reference's
examples): multi-restart fit, update, EI / LogEI / WIPV next point, kriging-believer batch, single-point calls."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP, EI, LogEI, WIPV, SurrogatePool
from oracle import gp_oracle as O

def timed(label, fn, reps=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print(f"{label:55s} {dt*1e3:9.2f} ms")
    return out

for n, d in ((200, 2), (500, 6), (1500, 12)):
    print(f"--- n={n} d={d}")
    X, y = O.synthetic_training_set(n, d)
    gp = GP(X, y, kernel="rbf", lengthscales=np.full(d, 0.5))
    ref = O.OracleGP(X, y, kernel="rbf", lengthscales=np.full(d, 0.5))
    x0 = O.synthetic_restarts(ref, 8)
    gp.fit(x0[:2], maxiter=2)
    res = timed("fit: 8 restarts, L-BFGS-B maxiter=50", lambda: gp.fit(x0, maxiter=50))
    gp.update_hyperparams(res['params'])
    rng = np.random.default_rng(0)
    timed("update: +1 point (incremental)", lambda: gp.update(rng.uniform(0, 1, (1, d)), np.array([[-1.0]])), reps=3)
    timed("predict_mean_single", lambda: gp.predict_mean_single(rng.uniform(0, 1, d)), reps=20)
    timed("predict_single (mean+var)", lambda: gp.predict_single(rng.uniform(0, 1, d)), reps=20)
    best = float(gp.train_y.max())
    for cls in (EI, LogEI):  # wall time AND the number of batched value-and-gradient calls / points the optimiser issued
        acq = cls()
        cnt = {"calls": 0, "pts": 0}
        orig = acq.value_and_grad_batched
        def counted(xs, *a, _o=orig, **k):
            cnt["calls"] += 1; cnt["pts"] += len(np.atleast_2d(xs))
            return _o(xs, *a, **k)
        acq.value_and_grad_batched = counted
        timed(f"{cls.__name__}.get_next_point (20 restarts, maxiter 250)", lambda: acq.get_next_point(gp, {'best_y': best, 'zeta': 0.01}, verbose=False, rng=rng))
        print(f"    -> {cnt['calls']} batched value+gradient calls, {cnt['pts']} point evaluations")
    mc = {'x': rng.uniform(0, 1, (2048, d))}
    timed("WIPV.get_next_point (mc_points_size=256)", lambda: WIPV().get_next_point(gp, {'mc_samples': mc, 'mc_points_size': 256}, verbose=False, rng=rng))
    timed("WIPV.get_next_batch (n_batch=4)", lambda: WIPV().get_next_batch(gp, n_batch=4, acq_kwargs={'mc_samples': mc, 'mc_points_size': 256}, verbose=False, rng=rng))
    pool = SurrogatePool(gp, size=64)
    pts = rng.uniform(0, 1, (64, 50, d))
    timed("SurrogatePool: 64 walks x 50 single-point loglike", lambda: pool.map(lambda i: [pool.loglike(p) for p in pts[i]], range(64)))
