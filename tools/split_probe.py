"""Predict mean+var at the headline shape (n = 2000, d = 16, Matern) over 8 full chunks; BOBE_TRMM_SPLIT / BOBE_KCHUNKS
select the schedule.  Prints ms per call (CUDA events) and a checksum of the variances."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP
from oracle import gp_oracle as O
X, y = O.synthetic_training_set(2000, 16)
gp = GP(X, y, kernel="matern", lengthscales=np.ones(16))
M = 148 * 128 * 8
Xq = torch.as_tensor(O.synthetic_queries(M, 16), device="cuda")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(2): m, v = gp.predict_mean_var_batched(Xq)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): m, v = gp.predict_mean_var_batched(Xq)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters
print(f"split={os.environ.get('BOBE_TRMM_SPLIT','1')} kchunks={os.environ.get('BOBE_KCHUNKS','-')}: {t:.3f} ms per {M} queries -> {M/t*1e3:.4e} pts/s; var sum {v.sum().item():.12e} max {v.max().item():.6e}")
