"""WIPV config E timing (development aid): n=4000, d=12, n_mc=1e5, C=8."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import GP
from oracle import gp_oracle as O
n, d, n_mc, C = 4000, 12, 100_000, 8
X, y = O.synthetic_training_set(n, d)
gp = GP(X, y, kernel="rbf", lengthscales=np.full(d, 1.0))
mc = torch.as_tensor(O.synthetic_queries(n_mc, d), device="cuda")
cand = torch.as_tensor(O.synthetic_queries(C, d, seed=6), device="cuda")
for _ in range(2): gp.fantasy_acquisition(mc, cand)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): out = gp.fantasy_acquisition(mc, cand)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 3
fl = n * n_mc * (3 * d + 3) + n * n * (n_mc + C) + 2 * n * C * n_mc + 4 * C * n_mc
print(f"WIPV E: {t:.2f} ms, {fl/t/1e9:.2f} TF algorithmic = {fl/t/1e9/35.46*100:.1f} % of DGEMM; out[0]={out[0].item():.12e}")
