"""Quick device-side timing of the two headline kernels (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import ops
torch.manual_seed(0)
dev = "cuda"
def ev_time(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
n, d = 2000, 16
X = torch.rand(n, d, dtype=torch.float64, device=dev)
y = (-0.5 * (((X - 0.5) / 0.15) ** 2).sum(1)); y = (y - y.mean()) / y.std()
ls = torch.ones(d, dtype=torch.float64, device=dev)
t = ev_time(lambda: ops.factorize("matern", X, y, ls[None], torch.ones(1, dtype=torch.float64, device=dev), 1e-8), iters=3)
print(f"factorize n={n}: {t:.3f} ms  ({(2*n**3/3)/t/1e9:.2f} TFLOP/s on chol+inv)")
L, Linv, alpha, logdet, quad, info = ops.factorize("matern", X, y, ls[None], torch.ones(1, dtype=torch.float64, device=dev), 1e-8)
for M in (18944, 148*128*8):
    Xq = torch.rand(M, d, dtype=torch.float64, device=dev)
    t = ev_time(lambda: ops.predict("matern", X, ls, 1.0, 1e-8, Linv[0], alpha[0], Xq, 0.0, 1.0), iters=2)
    fl = M * (n * n + n * (3 * d + 10) + 4 * n)
    print(f"predict mean+var M={M}: {t:.3f} ms -> {M/t*1e3:.3e} pts/s, {fl/t/1e9:.2f} TFLOP/s algorithmic")
    t = ev_time(lambda: ops.predict("matern", X, ls, 1.0, 1e-8, Linv[0], alpha[0], Xq, 0.0, 1.0, want_var=False), iters=2)
    print(f"predict mean only M={M}: {t:.3f} ms -> {M/t*1e3:.3e} pts/s")
for R in (8, 64):
    lp = torch.log(torch.cat([torch.ones(R, d, dtype=torch.float64, device=dev) * (0.5 + torch.rand(R, d, dtype=torch.float64, device=dev)), torch.ones(R, 1, dtype=torch.float64, device=dev)], 1))
    t = ev_time(lambda: ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8), iters=2)
    fl = R * (n ** 3 + n * n * (5 * d + 10 + 8))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8)
    t_enq = (time.perf_counter() - t0) / 5 * 1e3; torch.cuda.synchronize()
    print(f"mll+grad R={R}: {t:.3f} ms -> {R/t*1e3:.1f} evals/s, {fl/t/1e9:.2f} TFLOP/s algorithmic (host enqueue {t_enq:.3f} ms)")

# CUDA-graph replay of the whole batched log-ML+grad call (the C-ABI is capture-safe: no sync, no allocation)
for R in (1, 8, 64):
    lp = torch.log(torch.cat([torch.ones(R, d, dtype=torch.float64, device=dev) * (0.5 + torch.rand(R, d, dtype=torch.float64, device=dev)), torch.ones(R, 1, dtype=torch.float64, device=dev)], 1))
    ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            out = ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8)
    torch.cuda.synchronize()
    t = ev_time(lambda: g.replay(), iters=5)
    t2 = ev_time(lambda: ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8), iters=5)
    print(f"R={R}: graph replay {t:.3f} ms vs eager {t2:.3f} ms; val[0] {out[0][0].item():.6f}")
