#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "predict or golden or small_training or edge or kernel_variants or shared_panel or config_d or headline" > gpurun_out/r02_split_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_split_tests.log
python tools/timeline.py predict 500 4 100000 rbf 2>&1 | grep -v Warn | tail -4
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from bobe_b200 import GP
from oracle import gp_oracle as O
def ev_time(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for n, d, M in ((500, 4, 100000), (2000, 16, 2048), (2000, 16, 6000), (1000, 8, 4096)):
    X, y = O.synthetic_training_set(n, d)
    gp = GP(X, y, kernel="rbf", lengthscales=np.full(d, 0.5), device="cuda")
    Xq = torch.as_tensor(O.synthetic_queries(M, d), device="cuda")
    t = ev_time(lambda: gp.predict_mean_var_batched(Xq))
    print(f"n={n} d={d} M={M}: {t:.3f} ms  {M / t * 1e3:.3e} pts/s  {M * n * n / t / 1e9 / 35.46:.3f} of peak")
PY
