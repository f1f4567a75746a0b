"""Sub-batch policy sweep: log-ML+grad round for several R with the stream count forced by the environment (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bobe_b200 import ops
from tools.factor_ab import ev_time, data
X, y = data(2000, 16)
out = []
for R in (8, 12, 16, 24, 32, 48, 64):
    lp = torch.log(torch.cat([torch.ones(R, 16, dtype=torch.float64, device="cuda") * (0.5 + torch.rand(R, 16, dtype=torch.float64, device="cuda")),
                              torch.ones(R, 1, dtype=torch.float64, device="cuda")], 1))
    t = ev_time(lambda: ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8), iters=5, warm=2)
    out.append(f"R={R}: {t:.2f}")
print(os.environ.get("BOBE_MLL_STREAMS"), os.environ.get("BOBE_LOOKAHEAD_MAX", "8"), " ".join(out))
