"""Time the log-ML+grad round at R = 64 / 16 / 8 only (development aid, knobs via environment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bobe_b200 import ops
from tools.factor_ab import ev_time, data
n, d = 2000, 16
X, y = data(n, d)
out = []
for R in (64, 16, 8):
    lp = torch.log(torch.cat([torch.ones(R, d, dtype=torch.float64, device="cuda") * (0.5 + torch.rand(R, d, dtype=torch.float64, device="cuda")),
                              torch.ones(R, 1, dtype=torch.float64, device="cuda")], 1))
    t = ev_time(lambda: ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8), iters=6, warm=2)
    tg = ev_time(lambda: ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8, reuse_buffers=True), iters=6, warm=3)
    import time
    def sync_call():
        v, g, i = ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8, reuse_buffers=True)
        return v.cpu()
    for _ in range(3): sync_call()
    t0 = time.perf_counter()
    for _ in range(10): sync_call()
    ts = (time.perf_counter() - t0) / 10 * 1e3
    def sync_call_eager():
        v, g, i = ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8)
        return v.cpu()
    for _ in range(3): sync_call_eager()
    t0 = time.perf_counter()
    for _ in range(10): sync_call_eager()
    te = (time.perf_counter() - t0) / 10 * 1e3
    out.append(f"R={R}: back-to-back {t:.3f} ms, graph {tg:.3f} ms; call+fetch wall: graph {ts:.3f} ms, plain {te:.3f} ms")
print({k: v for k, v in os.environ.items() if k.startswith("BOBE_")}, " | ".join(out))
