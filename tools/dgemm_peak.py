"""Measure the cuBLAS DGEMM peak (FP64 roofline denominator) the same way MEASURED_PEAKS.json measures bf16."""
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3): torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best / 1e9
t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); k = 0
while time.time() - t0 < 4.0:
    for _ in range(5): torch.matmul(a, b, out=c)
    k += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = 2 * n**3 * k / e0.elapsed_time(e1) / 1e9
# HBM copy check
x = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize(); bb = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize(); bb = min(bb, e0.elapsed_time(e1))
print(json.dumps({"fp64_dgemm_tflops": burst, "fp64_dgemm_tflops_sustained": sus, "hbm_copy_gbs": 2 * x.numel() * 8 / bb / 1e6,
                  "how": "torch.matmul float64 8192^3 best of 10 (burst) and 4 s back to back (sustained)", "gpu": torch.cuda.get_device_name(0)}))
