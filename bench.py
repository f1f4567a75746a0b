#!/usr/bin/env python
"""Benchmark of the GP surrogate hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W          # our arm (under torchrun for N > 1)
  python bench.py --impl reference --gpus N ...          # the reference's CPU arithmetic (oracle port) on host cores

One JSON line.  Its top-level value is the headline metric of BASELINE.json -- posterior mean+variance evaluations/sec
at n=2000, d=16, Matern-5/2, M=10^6 queries PER GPU per step (weak scaling: queries shard with no data-path collective;
the closing all-gather of the 16 B/query results is inside the timed region for N > 1).  Beside it, from the same run:
  "strong"     the same sweep with M = 10^6 queries IN TOTAL split over the N GPUs (north_star's "M = 10^6 on 8 x B200");
  "secondary"  log-ML + gradient evals/sec over 64 restarts sharded over the ranks (strong), with its own clocks record;
  "wipv"       BASELINE config 5: WIPV over n_mc = 10^5 MC points x 8 candidates at n = 4000, d = 12, MC columns sharded.
For N > 1 every one of them goes through the product's sharding layer bobe_b200.dist (predict_sharded, mll_grad_sharded,
wipv_sharded), not through inline collectives.
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
M_STRONG = 1_000_000
R_TOTAL = 64
E_N, E_DIM, E_NMC, E_CAND = 4000, 12, 100_000, 8  # BASELINE config 5
CPU_CHUNK = 1024
WORKLOAD = "H: predict mean+var, n=2000 d=16 Matern-5/2 ARD, M=1e6 queries per GPU (synthetic, SURVEY.md 8d)"


def bench_config(world, m_per_gpu=M_PER_GPU):
    """The SAME dictionary in both arms (the reference arm times a bounded sample of this workload; what the sample was is
    said in its cpu_baseline.sample, not here)."""
    return {"workload": WORKLOAD, "n": N_TRAIN, "d": DIM, "kernel": KERNEL, "m_per_gpu": m_per_gpu,
            "l2": "inputs larger than L2 (128 MB of queries + 930 MB K* scratch per 3-chunk sweep)",
            "parallelism": f"query-sharded x{world}, replicated factor"}


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
        return self

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


# ---- the reference's CPU arithmetic (oracle port) on the host cores -------------------------------------------------------
# Thread policy, fixed explicitly so that the arm is the same at every N (torch.distributed.run exports OMP_NUM_THREADS=1,
# which silently halved this arm in round 1): the fastest of POLICIES (below) on a warm-up sample is used, with the BLAS
# thread count set through threadpoolctl rather than inherited from the environment.
def _blas_limits(n):
    from threadpoolctl import threadpool_limits
    return threadpool_limits(limits=int(n))


_CPU_GP = {}


def _cpu_gp():
    from oracle import gp_oracle as O
    if "gp" not in _CPU_GP:
        X, y = O.synthetic_training_set(N_TRAIN, DIM)
        _CPU_GP["gp"] = O.OracleGP(X, y, kernel=KERNEL, lengthscales=np.full(DIM, ELL))
    return _CPU_GP["gp"]


POLICIES = {  # name: (Python threads, BLAS threads per call, query chunk); None = all cores
    "threads": (None, 1, CPU_CHUNK),        # cores x 1: no oversubscription
    "blas": (1, None, 8 * CPU_CHUNK),       # 1 x cores
    "both": (None, None, CPU_CHUNK),        # cores x cores (round 1's arm; oversubscribed, but the NumPy kernel build holds
}                                           # the GIL between BLAS calls, so the extra BLAS threads fill those gaps)


def _policy_text(policy):
    cores = os.cpu_count()
    pt, bt, chunk = POLICIES[policy]
    return f"{pt or cores} Python threads x {bt or cores} BLAS threads, query chunks of {chunk}"


def cpu_predict_sample(n_queries, policy):
    """mean+var over a bounded query sample under one of the thread POLICIES."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import gp_oracle as O
    gp = _cpu_gp()
    Xq = O.synthetic_queries(n_queries, DIM)
    cores = os.cpu_count()
    pt, bt, chunk = POLICIES[policy]

    def work(s):  # query chunks keep the temporaries cache-sized; NumPy releases the GIL inside its loops
        gp.predict_mean_batched(Xq[s:s + chunk])
        gp.predict_var_batched(Xq[s:s + chunk])
    with _blas_limits(bt or cores):
        t0 = time.perf_counter()
        if (pt or cores) > 1:
            with ThreadPoolExecutor(pt or cores) as ex:
                list(ex.map(work, range(0, n_queries, chunk)))
        else:
            for s in range(0, n_queries, chunk):
                work(s)
        dt = time.perf_counter() - t0
    return n_queries / dt, dt


def cpu_pick_policy():
    best = None
    for policy in POLICIES:
        cpu_predict_sample(4096, policy)
        v, _ = cpu_predict_sample(16384, policy)
        if best is None or v > best[1]:
            best = (policy, v)
    return best[0]


def cpu_mll_grad_evals(n_evals):
    """value_and_grad(neg_mll) at n = 2000 (BOBE/gp.py:385-398 through optim.py:309), all-core BLAS; evals/s."""
    from oracle import gp_oracle as O
    gp = _cpu_gp()
    x0 = O.synthetic_restarts(gp, max(2, n_evals))
    with _blas_limits(os.cpu_count()):
        t0 = time.perf_counter()
        for r in range(n_evals):
            gp.neg_mll_and_grad(x0[r])
        dt = time.perf_counter() - t0
    return n_evals / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    policy = cpu_pick_policy()
    sample = 65536
    for _ in range(max(0, args.warmup - 1)):
        cpu_predict_sample(4096, policy)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_predict_sample(sample, policy)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = sample / (ms / 1e3)
    mll_v, mll_dt = cpu_mll_grad_evals(2)
    pol = _policy_text(policy)
    line = {"impl": "reference", "metric": "gp_predict_mean_var_pts_per_sec", "value": val, "unit": "pts/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": "pts/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} queries per step (of the 1e6-per-GPU workload), NumPy/SciPy(OpenBLAS) "
                                       f"restatement of BOBE/gp.py predict_mean+predict_var (JAX is not installable here); "
                                       f"{pol} (the fastest of {len(POLICIES)} thread policies on a warm-up sample; BLAS threads "
                                       f"set with threadpoolctl, so the arm is identical at every N)"},
            "e2e": {"value": val, "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "secondary": {"metric": "gp_mll_grad_evals_per_sec", "value": mll_v, "unit": "evals/s",
                          "sample": f"2 of the 64 restarts ({mll_dt:.1f} s), oracle neg_mll_and_grad at n={N_TRAIN}, "
                                    f"{cores} BLAS threads", "cores": cores, "kind": "port"}}
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
    ap.add_argument("--skip-extras", action="store_true", help="headline + roofline only (profiling runs)")
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
    from bobe_b200 import GP, ops, _lib, dist
    from oracle import gp_oracle as O  # synthetic input recipe + the cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG=VERSION (set in this image) makes NCCL print its version banner on STDOUT; stdout is already
        # redirected above, so nothing is silenced here: the setting the caller chose stays in force
        tdist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

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

    M = args.m_per_gpu
    M_all = M * world
    X, y = O.synthetic_training_set(N_TRAIN, DIM)
    gp = GP(X, y, kernel=KERNEL, lengthscales=np.full(DIM, ELL), kernel_variance=1.0, device=dev)
    lo, hi = dist.shard_bounds(M_all, rank, world)
    # the FULL query set is what the sharding layer takes; every rank only ever reads its own block, so only that block
    # is filled (device copy for the device-resident number, pinned host copy for the end-to-end number)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    Xq_all = torch.empty((M_all, DIM), dtype=torch.float64, device=dev)
    Xq_all[lo:hi] = torch.rand((hi - lo, DIM), dtype=torch.float64, device=dev, generator=gen)
    Xq_host_all = torch.empty((M_all, DIM), dtype=torch.float64, pin_memory=True) if world == 1 else None
    if world > 1:  # pin only this rank's block (N x 128 MB of pinned memory per rank otherwise)
        own = torch.empty((hi - lo, DIM), dtype=torch.float64, pin_memory=True)
        own.copy_(Xq_all[lo:hi])

        class _HostView:  # what dist.predict_sharded needs of the full host array: shape and the rank's own slice
            shape = (M_all, DIM)

            def __getitem__(self, s):
                assert s.start == lo and s.stop == hi
                return own
        Xq_host_all = _HostView()
    else:
        Xq_host_all.copy_(Xq_all)

    def step_device():  # device-resident queries, results gathered on every rank (16 B/query)
        return dist.predict_sharded(gp, Xq_all, want_var=True, gather=True)

    def step_e2e():  # public API, host buffers: H2D of the queries and D2H of mean/var inside the timed region
        # (each rank keeps its block of the results on its host: the all-gather is part of the device-resident number)
        return dist.predict_sharded(gp, Xq_host_all, want_var=True, gather=False)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(step_device, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    value = M_all / (ms_step / 1e3)

    ms_e2e = timed(step_e2e, max(2, min(args.steps, 3)), 1)
    e2e_value = M_all / (ms_e2e / 1e3)

    # ---- strong scaling of the same sweep: M = 1e6 queries in total, split over the ranks ------------------------------
    strong = None
    if not args.skip_extras:
        Ms = min(M_STRONG, M_all)
        Xq_s = Xq_all[:Ms] if world == 1 else torch.rand((Ms, DIM), dtype=torch.float64, device=dev,
                                                         generator=torch.Generator(device=dev).manual_seed(11))
        ms_s = timed(lambda: dist.predict_sharded(gp, Xq_s, want_var=True, gather=True), args.steps, 3)
        strong = {"metric": "gp_predict_mean_var_pts_per_sec", "value": Ms / (ms_s / 1e3), "unit": "pts/s",
                  "scaling": "strong", "m_total": Ms, "m_per_gpu": -(-Ms // world), "ms_per_step": ms_s, "steps": args.steps}

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
    # The FP64 roofline denominator, measured LIVE in this run the way MEASURED_PEAKS.json measures the bf16 one (which has
    # no FP64 entry): cuBLAS DGEMM 8192^3 through torch.matmul, best of 10, CUDA events.  A measurement tool only -- nothing
    # on the product path touches cuBLAS.  The file value (round 1, same recipe) is kept beside it.
    try:
        a8 = torch.rand((8192, 8192), dtype=torch.float64, device=dev)
        b8 = torch.rand((8192, 8192), dtype=torch.float64, device=dev)
        c8 = torch.empty_like(a8)
        best = None
        for i in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a8, b8, out=c8)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                t = e0.elapsed_time(e1)
                best = t if best is None else min(best, t)
        live = 2.0 * 8192.0 ** 3 / (best / 1e3) / 1e12
        peaks["fp64_dgemm_tflops_file"] = peaks["fp64_dgemm_tflops"]
        peaks["fp64_dgemm_tflops"] = live
        peaks["src"] = ("measured live in this run: cuBLAS DGEMM 8192^3 via torch.matmul, best of 10, CUDA events "
                        f"(FP64_PEAKS.json from round 1: {peaks['fp64_dgemm_tflops_file']:.2f}); MEASURED_PEAKS.json has no FP64 entry")
        del a8, b8, c8
    except Exception as exc:  # keep the file value
        peaks["src"] += f" (live DGEMM measurement failed: {exc})"
    achieved = flops_per_launch / (ms_k / 1e3) / 1e12
    chunks = -(-M // rows)
    traffic = None  # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (same chunk shape)
    for rnd in ("r02", "r01"):
        try:
            with open(os.path.join(ROOT, "profiles", rnd, "trmm_sumsq_ncu.json")) as f:
                j = json.load(f)
            traffic = float(j["dram_bytes_read"]) + float(j["dram_bytes_write"])
            break
        except Exception:
            pass
    roofline = {"bound": "tensor", "kernel": "trmm_sumsq_tma_kernel (FP64 DMMA.8x8x4 fed by TMA on mbarriers; no tcgen05 f64 kind exists)",
                "achieved": achieved, "peak": peaks["fp64_dgemm_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["fp64_dgemm_tflops"], "traffic": traffic,
                "peak_source": peaks["src"], "peak_file": peaks.get("fp64_dgemm_tflops_file"),
                "flops_per_launch": flops_per_launch, "ms_per_launch": ms_k,
                "share_of_step": chunks * ms_k / ms_step}
    del kstar

    # ---- secondary BASELINE metric: log-ML + gradient evals/sec, 64 restarts sharded over the ranks (strong) -----------
    secondary = wipv = None
    if not args.skip_extras:
        ref_gp = O.OracleGP(X, y, kernel=KERNEL, lengthscales=np.full(DIM, ELL))
        x0 = O.synthetic_restarts(ref_gp, R_TOTAL)
        s2 = ClockSampler(local_rank)
        if rank == 0:
            s2.start()
        ms_mll = timed(lambda: dist.mll_grad_sharded(gp, x0), max(10, args.steps), 3)
        clocks2 = s2.stop() if rank == 0 else None
        v_all, _ = dist.mll_grad_sharded(gp, x0)
        mll_flops = R_TOTAL * (N_TRAIN ** 3 + N_TRAIN ** 2 * (5 * DIM + 10 + 8))
        r_lo, r_hi = dist.shard_bounds(R_TOTAL, rank, world)
        secondary = {"metric": "gp_mll_grad_evals_per_sec", "value": R_TOTAL / (ms_mll / 1e3), "unit": "evals/s",
                     "restarts_total": R_TOTAL, "restarts_per_gpu": r_hi - r_lo, "ms_per_round": ms_mll,
                     "steps": max(10, args.steps), "warmup": 3, "non_pd_restarts": int(np.isnan(v_all).sum()),
                     "scaling": "strong", "api": "bobe_b200.dist.mll_grad_sharded -> GP.neg_mll_and_grad_batched (host in/out, "
                     "priors on the host, all-gather of (value, gradient) rows inside the timed region)",
                     "algorithmic_tflops": mll_flops / (ms_mll / 1e3) / 1e12,
                     "frac_of_fp64_peak": mll_flops / (ms_mll / 1e3) / 1e12 / (peaks["fp64_dgemm_tflops"] * world),
                     "clocks": clocks2}

        # ---- BASELINE config 5: WIPV, n = 4000, d = 12, n_mc = 1e5, C = 8, MC columns sharded (strong) ----------------
        Xe, ye = O.synthetic_training_set(E_N, E_DIM)
        gpe = GP(Xe, ye, kernel="rbf", lengthscales=np.full(E_DIM, 1.0), kernel_variance=1.0, device=dev)
        mc = torch.rand((E_NMC, E_DIM), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        cand = torch.rand((E_CAND, E_DIM), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(6))
        ms_w = timed(lambda: dist.wipv_sharded(gpe, mc, cand), max(5, args.steps), 2)
        w_flops = float(E_N) ** 2 * (E_NMC + E_CAND) + E_N * E_NMC * (3 * E_DIM + 3) + 2.0 * E_N * E_CAND * E_NMC
        wipv = {"metric": "wipv_acquisitions_per_sec", "value": 1e3 / ms_w, "unit": "calls/s", "ms_per_call": ms_w,
                "scaling": "strong", "n": E_N, "d": E_DIM, "n_mc": E_NMC, "candidates": E_CAND,
                "mc_points_per_sec": E_NMC / (ms_w / 1e3), "api": "bobe_b200.dist.wipv_sharded (MC columns sharded)",
                "algorithmic_tflops": w_flops / (ms_w / 1e3) / 1e12,
                "frac_of_fp64_peak": w_flops / (ms_w / 1e3) / 1e12 / (peaks["fp64_dgemm_tflops"] * world)}
        del gpe, mc

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        policy = cpu_pick_policy()
        sample = 131072
        v, dt = cpu_predict_sample(sample, policy)
        cores = os.cpu_count()
        pol = _policy_text(policy)
        cpu_baseline = {"value": v, "unit": "pts/s", "cores": cores, "kind": "port",
                        "sample": f"{sample} of the 1e6 queries ({dt:.1f} s), NumPy/SciPy(OpenBLAS) restatement of "
                                  f"BOBE/gp.py predict_mean+predict_var, {pol}; JAX is not installable here"}
        if secondary is not None:
            mv, mdt = cpu_mll_grad_evals(2)
            secondary["cpu_baseline"] = {"value": mv, "unit": "evals/s", "cores": cores, "kind": "port",
                                         "sample": f"2 of the 64 restarts ({mdt:.1f} s): oracle neg_mll_and_grad at n={N_TRAIN}, "
                                                   f"{cores} BLAS threads"}

    if rank == 0:
        line = {"metric": "gp_predict_mean_var_pts_per_sec", "value": value, "unit": "pts/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": bench_config(world, M),
                "e2e": {"value": e2e_value, "unit": "pts/s", "h2d_bytes_per_step": (hi - lo) * DIM * 8,
                        "d2h_bytes_per_step": (hi - lo) * 16, "ms_per_step": ms_e2e,
                        "api": "bobe_b200.dist.predict_sharded(gp, host queries) -> GP._predict (pinned staging, pipelined)"},
                "gpu_launches": args.steps * (chunks + -(-chunks // 3) + 1),  # per step: prescale + trmm_sumsq per chunk + kmat per 3 chunks
                "clocks": clocks, "roofline": roofline}
        if strong is not None:
            line["strong"] = strong
        if secondary is not None:
            line["secondary"] = secondary
        if wipv is not None:
            line["wipv"] = wipv
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
