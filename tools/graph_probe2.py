"""Sub-batch streams for few restarts (latency-bound chain): eager vs CUDA-graph replay at n = 2000."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bobe_b200 import ops
dev = "cuda"
def wall(fn, iters=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 2000, 16
for R in (4, 8, 16):
    X = torch.rand(n, d, dtype=torch.float64, device=dev)
    y = (-0.5 * (((X - 0.5) / 0.15) ** 2).sum(1)); y = (y - y.mean()) / y.std()
    lp = torch.log(torch.cat([0.5 + torch.rand(R, d, dtype=torch.float64, device=dev), torch.ones(R, 1, dtype=torch.float64, device=dev)], 1))
    f = lambda: ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8)
    t_eager = wall(f)
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        f(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            out = f()
    torch.cuda.synchronize()
    t_graph = wall(lambda: g.replay())
    print(f"n={n} R={R:2d} streams={os.environ.get('BOBE_MLL_STREAMS','4')} min/stream={os.environ.get('BOBE_MLL_MIN_PER_STREAM','4')}: eager {t_eager:7.3f} ms  graph {t_graph:7.3f} ms -> {R/t_graph*1e3:7.0f} evals/s (graph)")
