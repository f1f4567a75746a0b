"""Device-side timing of the kernel-matrix / predictive-mean kernel alone (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bobe_b200 import ops
dev = "cuda"
def ev_time(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (kind, n, d, c) in (("matern", 2000, 16, 10), ("rbf", 1500, 27, 3), ("rbf", 4000, 12, 3)):
    X = torch.rand(n, d, dtype=torch.float64, device=dev)
    ls = torch.ones(d, dtype=torch.float64, device=dev)
    alpha = torch.randn(ops.npad(n), dtype=torch.float64, device=dev)
    Linv = torch.eye(ops.npad(n), dtype=torch.float64, device=dev)
    M = 148 * 128 * 8
    Xq = torch.rand(M, d, dtype=torch.float64, device=dev)
    t = ev_time(lambda: ops.predict(kind, X, ls, 1.0, 1e-8, Linv, alpha, Xq, 0.0, 1.0, want_var=False))
    fl = M * n * (3 * d + c + 2)
    print(f"mean-only {kind} n={n} d={d} M={M}: {t:.3f} ms -> {M/t*1e3:.3e} pts/s, {fl/t/1e9:.2f} TF algorithmic, "
          f"{M*n*(2*d+ (33 if kind=='matern' else 20))/t/1e9/18.6e3*100:.1f}% of FP64 issue (est)")
    Ms = 18944
    t = ev_time(lambda: ops.kernel_matrix(kind, Xq[:Ms], X, ls, 1.0, 1e-8, False))
    print(f"   K*(store) rows={Ms}: {t:.3f} ms ({Ms*n*8/t/1e6:.0f} GB/s written)")
    t = ev_time(lambda: ops.kernel_matrix(kind, X, X, ls, 1.0, 1e-8, True))
    print(f"   K(X,X): {t:.3f} ms")
