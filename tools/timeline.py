"""Kernel timeline (CUPTI through torch.profiler) of one factorisation / one log-ML+grad round: start offset, duration,
stream and name of every kernel, so that gaps on the critical path and the overlap between streams can be read off.
  python tools/timeline.py factor [n]   |   python tools/timeline.py mll R [n]   |   python tools/timeline.py predict n d M [kernel]
"""
import os, sys, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bobe_b200 import ops

what = sys.argv[1] if len(sys.argv) > 1 else "factor"
dev = "cuda"
torch.manual_seed(0)
if what == "factor":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    R = 1
else:
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
d = 16
X = torch.rand(n, d, dtype=torch.float64, device=dev)
y = (-0.5 * (((X - 0.5) / 0.15) ** 2).sum(1)); y = (y - y.mean()) / y.std()
ls = torch.ones(1, d, dtype=torch.float64, device=dev)
kv = torch.ones(1, dtype=torch.float64, device=dev)
lp = torch.log(torch.cat([torch.ones(R, d, dtype=torch.float64, device=dev) * (0.5 + torch.rand(R, d, dtype=torch.float64, device=dev)),
                          torch.ones(R, 1, dtype=torch.float64, device=dev)], 1))
fn = (lambda: ops.factorize("matern", X, y, ls, kv, 1e-8)) if what == "factor" else \
     (lambda: ops.mll_grad_batched("matern", X, y, lp, True, 1.0, 1e-8))
if what == "predict":  # python tools/timeline.py predict n d M [kernel]
    import numpy as np
    from bobe_b200 import GP
    from oracle import gp_oracle as O
    n, d, M = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    kern = sys.argv[5] if len(sys.argv) > 5 else "rbf"
    Xh, yh = O.synthetic_training_set(n, d)
    gp = GP(Xh, yh, kernel=kern, lengthscales=np.full(d, 0.3 if kern == "rbf" else 1.0), device=dev)
    Xq = torch.as_tensor(O.synthetic_queries(M, d), device=dev)
    fn = lambda: gp.predict_mean_var_batched(Xq)  # noqa: E731
    R = 0
for _ in range(3):
    fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fn()
    torch.cuda.synchronize()
import json, tempfile
tmp = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tmp)
tr = json.load(open(tmp))
evs = [e for e in tr["traceEvents"] if e.get("cat") == "kernel"]
evs.sort(key=lambda e: e["ts"])
t0 = evs[0]["ts"]
tend = max(e["ts"] + e["dur"] for e in evs)
print(f"# {what} n={n} R={R}: {len(evs)} kernels, span {(tend - t0):.1f} us, sum of durations {sum(e['dur'] for e in evs):.1f} us")
streams = {}
prev_end = {}
busy = []
for e in evs:
    name = re.sub(r"\(.*", "", e["name"]).replace("void ", "").replace("bobe::", "")
    name = re.sub(r"TileCfg<([^>]*)>", lambda m: "T<" + m.group(1).replace(" ", "") + ">", name)[:58]
    s = e.get("args", {}).get("stream", -1)
    sid = streams.setdefault(s, len(streams))
    st, en = e["ts"] - t0, e["ts"] + e["dur"] - t0
    gap = st - prev_end.get(sid, st)
    prev_end[sid] = en
    g = e.get("args", {}).get("grid", "")
    print(f"{st:9.1f} {en - st:8.1f} us  s{sid}  gap {gap:7.1f}  {name}  grid={g}")
