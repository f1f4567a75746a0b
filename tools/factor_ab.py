"""Correctness + timing of the factorisation scheme selected by the BOBE_FACTOR / BOBE_FACTOR_PW / BOBE_LOOKAHEAD_MAX
knobs (development aid; the knobs are read once per process, so run one process per setting).

  python tools/factor_ab.py check      residuals of L, Linv, alpha, logdet against torch float64 (cuSOLVER) for many n
  python tools/factor_ab.py time       factorise one n = 2000 matrix; log-ML+grad rounds R = 8 / 16 / 64
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bobe_b200 import ops

dev = "cuda"
torch.manual_seed(0)


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def data(n, d):
    X = torch.rand(n, d, dtype=torch.float64, device=dev)
    y = (-0.5 * (((X - 0.5) / 0.15) ** 2).sum(1))
    y = (y - y.mean()) / y.std()
    return X, y


def check():
    worst = 0.0
    for n, d, kind, ell, B in [(2, 2, "rbf", 0.5, 1), (64, 2, "rbf", 0.5, 2), (65, 3, "matern", 0.7, 1), (100, 2, "rbf", 0.3, 1),
                               (128, 4, "matern", 0.7, 3), (129, 4, "matern", 0.7, 1), (192, 4, "rbf", 0.8, 2),
                               (300, 3, "matern", 0.7, 2), (320, 3, "matern", 0.7, 1), (500, 6, "rbf", 0.5, 2),
                               (500, 2, "rbf", 0.3, 1), (1000, 8, "matern", 1.0, 2), (1088, 8, "matern", 1.0, 1),
                               (1500, 27, "rbf", 2.0, 1), (2000, 16, "matern", 1.0, 2), (2100, 16, "matern", 1.0, 1),
                               (4000, 12, "rbf", 1.0, 1)]:
        X, y = data(n, d)
        ls = torch.full((B, d), ell, dtype=torch.float64, device=dev) * (1 + 0.1 * torch.arange(B, device=dev, dtype=torch.float64))[:, None]
        kv = torch.ones(B, dtype=torch.float64, device=dev)
        L, Linv, alpha, logdet, quad, info = ops.factorize(kind, X, y, ls, kv, 1e-8)
        torch.cuda.synchronize()
        for b in range(B):
            K = ops.kernel_matrix(kind, X, X, ls[b], 1.0, 1e-8, True)
            Lr = torch.linalg.cholesky(K)
            Lb, Xb = L[b, :n, :n], Linv[b, :n, :n]
            e_l = ((Lb @ Lb.T - K).abs().max() / K.abs().max()).item()
            e_i = ((Xb @ Lb - torch.eye(n, device=dev, dtype=torch.float64)).abs().max()).item()
            condL = (torch.linalg.norm(Lb, 2) * torch.linalg.norm(Xb, 2)).item()
            a_ref = torch.cholesky_solve(y[:, None], Lr)[:, 0]
            e_a = ((alpha[b, :n] - a_ref).abs().max() / a_ref.abs().max()).item()
            ld_ref = torch.log(torch.diagonal(Lr)).sum().item()
            e_d = abs(logdet[b].item() - ld_ref) / max(abs(ld_ref), n)
            up = max(torch.triu(L[b], 1).abs().max().item(), torch.triu(Linv[b], 1).abs().max().item())
            p = L.shape[1]
            pad_ok = True
            if p > n:
                pad_ok = bool(torch.equal(L[b, n:, n:], torch.eye(p - n, device=dev, dtype=torch.float64)) and L[b, n:, :n].abs().max().item() == 0.0)
            print(f"n={n:5d} d={d:2d} {kind:6s} b={b} info={int(info[b])} |LL^T-K|={e_l:.1e} |XL-I|={e_i:.1e} (cond L {condL:.1e}) "
                  f"alpha {e_a:.1e} logdet {e_d:.1e} upper {up:.1e} pad {pad_ok}")
            worst = max(worst, e_l, e_d)
            assert int(info[b]) == 0 and e_l < 1e-13 and e_d < 1e-9 and up == 0.0 and pad_ok and e_i < 1e-15 * condL * n, "FAILED"
    # non-PD input: NaN + info, never an error
    X, y = data(200, 2)
    X[1] = X[0]
    ls = torch.full((1, 2), 0.5, dtype=torch.float64, device=dev)
    L, Linv, alpha, logdet, quad, info = ops.factorize("rbf", X, y, ls, torch.ones(1, dtype=torch.float64, device=dev), -1e-6)
    print("non-PD: info", int(info[0]), "logdet", logdet[0].item())
    assert int(info[0]) == 1 and not torch.isfinite(logdet[0])
    print("check ok, worst", worst)


def timing():
    n, d = 2000, 16
    X, y = data(n, d)
    ls = torch.ones(1, d, dtype=torch.float64, device=dev)
    kv = torch.ones(1, dtype=torch.float64, device=dev)
    t = ev_time(lambda: ops.factorize("matern", X, y, ls, kv, 1e-8), iters=10, warm=3)
    print(f"factorize n={n}: {t:.3f} ms")
    for nn, dd in [(500, 6), (1500, 12), (4000, 12)]:
        X2, y2 = data(nn, dd)
        ls2 = torch.ones(1, dd, dtype=torch.float64, device=dev)
        t = ev_time(lambda: ops.factorize("matern", X2, y2, ls2, kv, 1e-8), iters=5, warm=2)
        print(f"factorize n={nn}: {t:.3f} ms")
    for R in (1, 8, 16, 64):
        lp = torch.log(torch.cat([torch.ones(R, d, dtype=torch.float64, device=dev) * (0.5 + torch.rand(R, d, dtype=torch.float64, device=dev)),
                                  torch.ones(R, 1, dtype=torch.float64, device=dev)], 1))
        t = ev_time(lambda: ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8), iters=5 if R < 64 else 3, warm=2)
        fl = R * (n ** 3 + n * n * (5 * d + 10 + 8))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8)
        t_enq = (time.perf_counter() - t0) / 3 * 1e3
        torch.cuda.synchronize()
        print(f"mll+grad R={R}: {t:.3f} ms -> {R / t * 1e3:.1f} evals/s, {fl / t / 1e9:.2f} TFLOP/s algorithmic (host enqueue {t_enq:.3f} ms)")
    X5, y5 = data(500, 6)
    lp = torch.log(torch.cat([torch.ones(8, 6, dtype=torch.float64, device=dev) * 0.7, torch.ones(8, 1, dtype=torch.float64, device=dev)], 1))
    t = ev_time(lambda: ops.mll_grad_batched("matern", X5, y5, lp, True, 1.0, 1e-8), iters=10, warm=3)
    print(f"mll+grad n=500 R=8: {t:.3f} ms")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    print("env:", {k: v for k, v in os.environ.items() if k.startswith("BOBE_")})
    if mode == "check":
        check()
    else:
        timing()
