#!/usr/bin/env python
"""Benchmark of the GP surrogate hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W          # our arm (under torchrun for N > 1)
  python bench.py --impl reference --gpus N ...          # the reference's CPU arithmetic (oracle port) on host cores

Headline metric (BASELINE.json): posterior mean+variance evaluations/sec at n=2000, d=16, Matern-5/2,
M=10^6 queries per GPU per step (weak scaling: queries shard with no data-path collective; the closing
all-gather of the 16 B/query results is inside the timed region for N > 1).  The second BASELINE metric,
log-ML+gradient evals/sec over 64 restarts (sharded across ranks), is reported in "secondary".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM, KERNEL, ELL = 2000, 16, "matern", 1.0
M_PER_GPU = 1_000_000
R_TOTAL = 64
CPU_CHUNK = 1024
WORKLOAD = "H: predict mean+var, n=2000 d=16 Matern-5/2 ARD, M=1e6 queries per GPU (synthetic, SURVEY.md 8d)"


def _peaks():
    p = {"fp64_dgemm_tflops": 35.46, "hbm_gbs": 6467.7, "src": "fallback constants"}
    try:
        with open(os.path.join(ROOT, "FP64_PEAKS.json")) as f:
            j = json.load(f)
        p["fp64_dgemm_tflops"] = float(j["fp64_dgemm_tflops"])
        p["src"] = "FP64_PEAKS.json (cuBLAS DGEMM 8192^3 measured on this pool's B200, round 1; MEASURED_PEAKS.json has no FP64 entry)"
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p["hbm_gbs"] = float(json.load(f)["hbm_gbs"])
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def cpu_predict_sample(n_queries, threads=None):
    """The oracle (NumPy/SciPy port of BOBE/gp.py) on host cores: mean+var over a bounded query sample."""
    from oracle import gp_oracle as O
    X, y = O.synthetic_training_set(N_TRAIN, DIM)
    gp = O.OracleGP(X, y, kernel=KERNEL, lengthscales=np.full(DIM, ELL))
    from concurrent.futures import ThreadPoolExecutor
    Xq = O.synthetic_queries(n_queries, DIM)
    threads = threads or os.cpu_count()

    def work(s):  # query chunks keep the temporaries cache-sized; NumPy releases the GIL inside its loops
        gp.predict_mean_batched(Xq[s:s + CPU_CHUNK])
        gp.predict_var_batched(Xq[s:s + CPU_CHUNK])
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(0, n_queries, CPU_CHUNK)))
    dt = time.perf_counter() - t0
    return n_queries / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    sample = 65536
    for _ in range(args.warmup):
        cpu_predict_sample(2048)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_predict_sample(sample)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": "gp_predict_mean_var_pts_per_sec", "value": val, "unit": "pts/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": N_TRAIN, "d": DIM, "kernel": KERNEL},
            "cpu_baseline": {"value": val, "unit": "pts/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} queries per step (of the 1e6 workload), NumPy/SciPy OpenBLAS restatement "
                                       f"of BOBE/gp.py (JAX is not installable here), {cores} threads x query chunks of {CPU_CHUNK}"},
            "e2e": {"value": val, "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _ensure_native():
    """The in-tree .so normally travels with the repo snapshot; if it is missing, local rank 0 builds it (nvcc) and
    the other ranks wait for the file.  There is no fallback: without the library the import below fails loudly."""
    LIB_PATH = os.path.join(ROOT, "bobe_b200", "lib", "libbobe_b200.so")
    if os.path.exists(LIB_PATH):
        return
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        import importlib.util  # by path: importing bobe_b200.build would import the package, which needs the library
        spec = importlib.util.spec_from_file_location("bobe_b200_build", os.path.join(ROOT, "bobe_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build_native(force=False)
    else:
        t0 = time.time()
        while not os.path.exists(LIB_PATH) and time.time() - t0 < 900:
            time.sleep(2.0)
        time.sleep(2.0)  # let the linker finish writing


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m-per-gpu", type=int, default=M_PER_GPU)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout must carry exactly ONE line (the JSON): native libraries (e.g. NCCL's version banner) write to fd 1, so
    # fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    _ensure_native()
    import torch
    import torch.distributed as tdist
    from bobe_b200 import GP, ops, _lib
    from oracle import gp_oracle as O  # synthetic input recipe + the cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG=VERSION (set in this image) makes NCCL print its version banner on STDOUT, in front of the one
        # JSON line the driver parses; keep warnings, drop the banner (an explicit INFO / TRACE request is respected)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        tdist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    M = args.m_per_gpu
    X, y = O.synthetic_training_set(N_TRAIN, DIM)
    gp = GP(X, y, kernel=KERNEL, lengthscales=np.full(DIM, ELL), kernel_variance=1.0, device=dev)
    rngq = np.random.default_rng(1 + rank)
    Xq_host = torch.from_numpy(rngq.uniform(0.0, 1.0, (M, DIM))).pin_memory()
    Xq = Xq_host.to(dev)
    gathered = [torch.empty(2 * M, dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def step_device():
        mean, var = gp.predict_mean_var_batched(Xq)
        if world > 1:  # closing all-gather of the sharded results (16 B/query)
            tdist.all_gather(gathered, torch.cat([mean, var]))
        return mean, var

    def step_e2e():  # public API, host buffers: H2D of the queries and D2H of mean/var inside the timed region
        mean, var = gp.predict_mean_var_batched(Xq_host)
        return mean, var

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        return float(ms.item()) / steps

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(step_device, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    value = world * M / (ms_step / 1e3)

    ms_e2e = timed(step_e2e, max(2, min(args.steps, 3)), 1)
    e2e_value = world * M / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel (trmm_sumsq), timed alone with CUDA events on the launching stream ----
    npad = ops.npad(N_TRAIN)
    rows = 148 * 128
    kstar = torch.rand((rows, npad), dtype=torch.float64, device=dev)
    kstar[:, N_TRAIN:] = 0.0
    vout = torch.empty(rows, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def launch_trmm():
        _lib.check(_lib.lib.bobe_bench_trmm_sumsq(stream, gp._Linv_dev.data_ptr(), N_TRAIN, kstar.data_ptr(), rows,
                                                  1.0, vout.data_ptr()), "bobe_bench_trmm_sumsq")
    ms_k = timed(launch_trmm, 10, 3)
    flops_per_launch = float(rows) * float(N_TRAIN) ** 2  # n^2 flops per query for the triangular apply
    peaks = _peaks()
    achieved = flops_per_launch / (ms_k / 1e3) / 1e12
    chunks = -(-M // rows)
    traffic = None  # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (same chunk shape)
    try:
        with open(os.path.join(ROOT, "profiles", "r01", "trmm_sumsq_ncu.json")) as f:
            j = json.load(f)
        traffic = float(j["dram_bytes_read"]) + float(j["dram_bytes_write"])
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "trmm_sumsq_tma_kernel (FP64 DMMA.8x8x4 fed by TMA on mbarriers; no tcgen05 f64 kind exists)",
                "achieved": achieved, "peak": peaks["fp64_dgemm_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["fp64_dgemm_tflops"], "traffic": traffic,
                "peak_source": peaks["src"], "flops_per_launch": flops_per_launch, "ms_per_launch": ms_k,
                "share_of_step": chunks * ms_k / ms_step}

    # ---- secondary BASELINE metric: log-ML + gradient evals/sec, 64 restarts sharded over the ranks -------------
    ref_gp = O.OracleGP(X, y, kernel=KERNEL, lengthscales=np.full(DIM, ELL))
    x0 = O.synthetic_restarts(ref_gp, R_TOTAL)
    lo, hi = rank * R_TOTAL // world, (rank + 1) * R_TOTAL // world
    lp = torch.as_tensor(x0[lo:hi], device=dev)
    allv = [torch.empty(hi - lo, dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def step_mll():
        val, grad, info = ops.mll_grad_batched(KERNEL, gp._X_dev, gp._y_dev, lp, True, 1.0, float(gp.noise))
        if world > 1:
            tdist.all_gather(allv, val)
        return val
    ms_mll = timed(step_mll, 3, 2)
    v_mll = step_mll()
    n_nan = int(torch.isnan(v_mll).sum().item())
    mll_flops = R_TOTAL * (N_TRAIN ** 3 + N_TRAIN ** 2 * (5 * DIM + 10 + 8))
    secondary = {"metric": "gp_mll_grad_evals_per_sec", "value": R_TOTAL / (ms_mll / 1e3), "unit": "evals/s",
                 "restarts_total": R_TOTAL, "restarts_per_gpu": hi - lo, "ms_per_round": ms_mll,
                 "non_pd_restarts_on_rank0": n_nan, "scaling": "strong",
                 "algorithmic_tflops": mll_flops / (ms_mll / 1e3) / 1e12,
                 "frac_of_fp64_peak": mll_flops / (ms_mll / 1e3) / 1e12 / (peaks["fp64_dgemm_tflops"] * world)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cpu_predict_sample(2048)
        sample = 131072
        v, dt = cpu_predict_sample(sample)
        cpu_baseline = {"value": v, "unit": "pts/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"{sample} of the 1e6 queries ({dt:.1f} s), NumPy/SciPy(OpenBLAS) restatement of "
                                  f"BOBE/gp.py predict_mean+predict_var, {os.cpu_count()} threads x query chunks of {CPU_CHUNK}; JAX is not installable here"}

    if rank == 0:
        line = {"metric": "gp_predict_mean_var_pts_per_sec", "value": value, "unit": "pts/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "n": N_TRAIN, "d": DIM, "kernel": KERNEL, "m_per_gpu": M,
                           "l2": "inputs larger than L2 (128 MB of queries + 930 MB K* scratch per 3-chunk sweep)",
                           "parallelism": f"query-sharded x{world}, replicated factor"},
                "e2e": {"value": e2e_value, "unit": "pts/s", "h2d_bytes_per_step": M * DIM * 8,
                        "d2h_bytes_per_step": M * 16, "ms_per_step": ms_e2e},
                "gpu_launches": args.steps * (chunks + -(-chunks // 3) + 1),  # per step: prescale + trmm_sumsq per chunk + kmat per 3 chunks
                "clocks": clocks, "roofline": roofline, "secondary": secondary}
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
