"""Eager vs CUDA-graph replay of one batched log-ML+grad call in the small-n (launch-bound) regime."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import ops
dev = "cuda"
def wall(fn, iters=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters): fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3
for n, d, R in ((100, 2, 8), (200, 2, 8), (500, 6, 8), (1000, 8, 8), (1500, 12, 8), (2000, 16, 8), (500, 6, 1), (500, 6, 32)):
    X = torch.rand(n, d, dtype=torch.float64, device=dev)
    y = (-0.5 * (((X - 0.5) / 0.15) ** 2).sum(1)); y = (y - y.mean()) / y.std()
    lp = torch.log(torch.cat([0.5 + torch.rand(R, d, dtype=torch.float64, device=dev), torch.ones(R, 1, dtype=torch.float64, device=dev)], 1))
    kern = "rbf"
    f = lambda: ops.mll_grad_batched(kern, X, y, lp, True, 1.0, 1e-8)
    t_eager = wall(f)
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        f(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            out = f()
    torch.cuda.synchronize()
    ref = f()
    g.replay(); torch.cuda.synchronize()
    same = torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    t_graph = wall(lambda: g.replay())
    print(f"n={n:5d} d={d:2d} R={R:2d}: eager {t_eager:7.3f} ms   graph {t_graph:7.3f} ms   x{t_eager/t_graph:4.2f}  bitwise {same}")
