"""Every BASELINE.json config on one B200: throughput + algorithmic work + fraction of its roofline.

Writes one JSON line per measurement to stdout (and profiles/r01/configs.jsonl when --out is given).  bench.py stays
the contract benchmark (config H + the 64-restart fit); this covers the other named shapes:
  A  Banana-like n=100, d=2, RBF: WIPV over n_mc=512 self-candidates (latency-bound: time only)
  B  n=500, d in {2,4,6}, RBF: predict mean+var M=1e5 (latency/launch-bound: time only)
  D  n=1500, d=27, RBF: mean-only sweep M=1e6 (FP64 pipe, kmat_kernel) and mean+var sweep
  E  n=4000, d=12, RBF: WIPV, n_mc=1e5 MC points x C=8 candidates
plus, for H, the "library bar": the same prediction through torch float64 (cuSOLVER/cuBLAS), i.e. what flipping the
reference to a GPU backend would roughly give.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bobe_b200 import GP, ops
from oracle import gp_oracle as O  # synthetic input recipe only

DGEMM_TF = 35.46  # FP64_PEAKS.json
FP64_ISSUE = 148 * 64 * 1.965e9  # FP64 lane-instructions/s (DFMA = 1), = 37.2 TFLOP/s nominal


def ev_time(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(out, **kw):
    line = json.dumps(kw)
    print(line, flush=True)
    if out:
        out.write(line + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-torch", action="store_true")
    args = ap.parse_args()
    out = open(args.out, "w") if args.out else None
    dev = torch.device("cuda", 0)

    # ---- A -------------------------------------------------------------------------------------------------------
    X, y = O.synthetic_training_set(100, 2)
    gp = GP(X, y, kernel="rbf", lengthscales=np.full(2, 0.3), device=dev)
    mc = torch.as_tensor(O.synthetic_queries(512, 2), device=dev)
    t = ev_time(lambda: gp.fantasy_acquisition(mc, None), iters=20, warm=3)
    emit(out, config="A", what="WIPV n=100 d=2 RBF, 512 MC points = candidates", ms=t, candidates_per_s=512 / t * 1e3,
         bound="latency (5 launches)")

    # ---- B -------------------------------------------------------------------------------------------------------
    for d in (2, 4, 6):
        X, y = O.synthetic_training_set(500, d)
        gp = GP(X, y, kernel="rbf", lengthscales=np.full(d, 0.3), device=dev)
        Xq = torch.as_tensor(O.synthetic_queries(100_000, d), device=dev)
        t = ev_time(lambda: gp.predict_mean_var_batched(Xq), iters=10, warm=2)
        fl = 1e5 * (500 ** 2 + 500 * (3 * d + 3) + 4 * 500)
        emit(out, config="B", what=f"predict mean+var n=500 d={d} RBF M=1e5", ms=t, pts_per_s=1e5 / t * 1e3,
             algorithmic_tflops=fl / t / 1e9, frac_of_dgemm_peak=fl / t / 1e9 / DGEMM_TF)
        x0 = torch.as_tensor(O.synthetic_restarts(O.OracleGP(X, y, kernel="rbf", lengthscales=np.full(d, 0.3)), 8), device=dev)
        t = ev_time(lambda: ops.mll_grad_batched("rbf", gp._X_dev, gp._y_dev, x0, True, 1.0, 1e-8), iters=5, warm=2)
        emit(out, config="B", what=f"log-ML+grad n=500 d={d} RBF R=8", ms=t, evals_per_s=8 / t * 1e3, bound="latency")

    # ---- D -------------------------------------------------------------------------------------------------------
    n, d, M = 1500, 27, 1_000_000
    X, y = O.synthetic_training_set(n, d)
    gp = GP(X, y, kernel="rbf", lengthscales=np.full(d, 2.0), device=dev)
    Xq = torch.as_tensor(O.synthetic_queries(M, d), device=dev)
    t = ev_time(lambda: gp.predict_mean_batched(Xq), iters=3, warm=1)
    fl = M * (n * (3 * d + 3) + 2 * n)
    lane = M * n * (2 * d + 20)  # FP64 instructions actually needed per element: 2 per dim + exp (18) + scale + dot
    emit(out, config="D", what="nested-sampling mean sweep n=1500 d=27 RBF M=1e6 (kmat_kernel, mean fused)", ms=t,
         pts_per_s=M / t * 1e3, algorithmic_tflops=fl / t / 1e9, bound="fp64 pipe (DFMA exp polynomial; 580 flop/B)",
         fp64_issue_frac=lane / (t / 1e3) / FP64_ISSUE, hbm_gbs=M * (d + 1) * 8 / t / 1e6)
    t = ev_time(lambda: gp.predict_mean_var_batched(Xq), iters=3, warm=1)
    fl = M * (n * n + n * (3 * d + 3) + 4 * n)
    emit(out, config="D", what="mean+var sweep n=1500 d=27 RBF M=1e6", ms=t, pts_per_s=M / t * 1e3,
         algorithmic_tflops=fl / t / 1e9, frac_of_dgemm_peak=fl / t / 1e9 / DGEMM_TF, bound="tensor (DMMA)")
    del Xq

    # ---- E -------------------------------------------------------------------------------------------------------
    n, d, n_mc, C = 4000, 12, 100_000, 8
    X, y = O.synthetic_training_set(n, d)
    gp = GP(X, y, kernel="rbf", lengthscales=np.full(d, 1.0), device=dev)
    mc = torch.as_tensor(O.synthetic_queries(n_mc, d), device=dev)
    cand = torch.as_tensor(O.synthetic_queries(C, d, seed=6), device=dev)
    t = ev_time(lambda: gp.fantasy_acquisition(mc, cand), iters=3, warm=1)
    fl = n * n_mc * (3 * d + 3) + n * n * (n_mc + C) + 2 * n * C * n_mc + 4 * C * n_mc
    emit(out, config="E", what="WIPV n=4000 d=12 RBF, n_mc=1e5 x C=8", ms=t, acq_values_per_s=C / t * 1e3,
         mc_points_per_s=n_mc / t * 1e3, algorithmic_tflops=fl / t / 1e9, frac_of_dgemm_peak=fl / t / 1e9 / DGEMM_TF,
         bound="tensor (DMMA): shared solve V = Linv K(X,MC)")
    tf = ev_time(lambda: ops.factorize("rbf", gp._X_dev, gp._y_dev, gp._ls_dev[None], torch.ones(1, dtype=torch.float64, device=dev), 1e-8), iters=3)
    emit(out, config="E", what="factorise n=4000 (K -> L, Linv, alpha), one matrix", ms=tf,
         algorithmic_tflops=(2 * n ** 3 / 3) / tf / 1e9, bound="latency chain (one matrix)")
    del mc

    # ---- H: library bar ------------------------------------------------------------------------------------------------
    if not args.skip_torch:
        n, d, M = 2000, 16, 151_552
        X, y = O.synthetic_training_set(n, d)
        gp = GP(X, y, kernel="matern", lengthscales=np.ones(d), device=dev)
        Xq = torch.as_tensor(O.synthetic_queries(M, d), device=dev)
        t_ours = ev_time(lambda: gp.predict_mean_var_batched(Xq), iters=3, warm=1)
        Xd = torch.as_tensor(X, device=dev)
        yd = gp._y_dev.reshape(-1, 1)

        def matern(a, b):
            q = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
            r = torch.sqrt(torch.clamp(q, min=1e-30))
            s5 = 5.0 ** 0.5
            return (1.0 + r * (s5 + r * (5.0 / 3.0))) * torch.exp(-s5 * r)
        K = matern(Xd, Xd) + 1e-8 * torch.eye(n, dtype=torch.float64, device=dev)
        L = torch.linalg.cholesky(K)
        alpha = torch.cholesky_solve(yd, L)

        def torch_predict():
            outs = []
            for s in range(0, M, 8192):  # chunked: the (chunk, n, d) broadcast temporaries stay within a few GB
                ks = matern(Xq[s:s + 8192], Xd)
                mean = ks @ alpha
                v = torch.linalg.solve_triangular(L, ks.T, upper=False)
                outs.append((mean, 1.0 + 1e-8 - (v * v).sum(0)))
            return outs
        t_lib = ev_time(torch_predict, iters=2, warm=1)
        emit(out, config="H", what="library bar: torch float64 (broadcast kernel + cuBLAS trsm) vs bobe_b200, M=151,552",
             ms_torch=t_lib, pts_per_s_torch=M / t_lib * 1e3, ms_ours=t_ours, pts_per_s_ours=M / t_ours * 1e3,
             speedup=t_lib / t_ours)
    if out:
        out.close()


if __name__ == "__main__":
    main()
