#!/usr/bin/env python
"""bench.py -- new_mpn_mul throughput on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3]

A "step" is one product of the workload (default BASELINE.json configs[1]: 2^20 x 2^20 limbs,
depth 14, w 1) on synthetic uniform limbs (splitmix64 counter generator).

  value      whole-job output limbs / s, operands and result resident in HBM, CUDA-event time per
             step (L2 flushed between steps), max over ranks
  e2e        the same metric through the reference-facing C symbol with HOST buffers (pinned),
             H2D/D2H inside the timed region
  roofline   the transform-stage kernel (k_run_stage): algorithmic bytes per launch / mean launch
             duration, from per-launch CUDA events in a separate profiling leg of the same run
  cpu_baseline  the compiled reference (oracle/_ref, mul_fft.c:3246 fixed) on one host core

--impl reference times only the reference's CPU implementation (all host cores, one product per
process per step).  N > 1: one process per GPU under torchrun; each rank multiplies its own
operands (independent products, weak scaling, no data-path collective).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (n1, n2, depth, w)  -- SURVEY section 8 table
    "cfg1": (1 << 16, 1 << 16, 12, 1),
    "cfg2": (1 << 20, 1 << 20, 14, 1),
    "cfg3": (3000000, 1700000, 14, 2),
}
METRIC, UNIT = "new_mpn_mul Mlimb/s", "Mlimb/s"


def splitmix64(seed, n):
    z = (np.arange(n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.proc = gpu, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        time.sleep(0.15)
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_product(n1, n2, depth, w, reps, seed=1):
    """time the compiled reference (fixed at mul_fft.c:3246) on this process's core"""
    from oracle import loader as oracle
    ref = oracle.load_ref(True)
    kind = "reference"
    if ref is None:
        ref = oracle.load_port()
        kind = "port"
    a, b = splitmix64(seed, n1), splitmix64(seed + 7, n2)
    r = np.zeros(n1 + n2, dtype=np.uint64)
    P = lambda x: C.c_void_p(x.ctypes.data)   # noqa: E731
    times = []
    for _ in range(reps):
        t = time.perf_counter()
        ref.new_mpn_mul(P(r), P(a), C.c_long(n1), P(b), C.c_long(n2), C.c_ulong(depth), C.c_ulong(w))
        times.append(time.perf_counter() - t)
    return times, kind


def _ref_worker(args):
    n1, n2, depth, w, reps, seed = args
    try:
        os.sched_setaffinity(0, {seed % os.cpu_count()})
    except Exception:
        pass
    return cpu_reference_product(n1, n2, depth, w, reps, seed)[0]


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU path, all host cores (independent products)"""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import loader as oracle
    n1, n2, depth, w = WORKLOADS[args.workload]
    kind = "reference" if oracle.load_ref(True) is not None else "port"
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_ref_worker, [(n1, n2, depth, w, max(1, min(args.warmup, 1)), s) for s in range(cores)])
        t0 = time.perf_counter()
        pool.map(_ref_worker, [(n1, n2, depth, w, args.steps, s) for s in range(cores)])
        dt = time.perf_counter() - t0
    products = cores * args.steps
    value = products * (n1 + n2) / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "new_mpn_mul %d x %d limbs, depth %d, w %d" % (n1, n2, depth, w),
                   "note": "each step = one product per host core, %d cores in parallel" % cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d products per core x %d cores, compiled reference mul_fft.c "
                                   "(3246 fixed) on GMP 6.3 shim" % (args.steps, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import mpir_fft_b200 as M

    if not torch.cuda.is_available() or not M.have_gpu():
        raise SystemExit("bench.py: no CUDA device -- mpir_fft_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    M.init(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n1, n2, depth, w = WORKLOADS[args.workload]
    L = M.lib()
    plan = M.MulPlan(n1, n2, depth, w)
    prm = plan.params
    S = 8 * (prm["limbs"] + 1)

    # operands resident in HBM (each rank its own product)
    a = torch.from_numpy(splitmix64(0x5EED0001 + 1000 * rank, n1).view(np.int64)).cuda()
    b = torch.from_numpy(splitmix64(0x5EED0002 + 1000 * rank, n2).view(np.int64)).cuda()
    r = torch.zeros(n1 + n2, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    pa, pb, pr = a.data_ptr(), b.data_ptr(), r.data_ptr()

    def step():
        plan.exec_device(pr, pa, pb, None)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    # correctness of what is being timed: compare with GMP on rank 0 (outside the timed region)
    check = None
    if rank == 0:
        try:
            from oracle import loader as oracle
            want = oracle.gmp_mul(a.cpu().numpy().view(np.uint64), b.cpu().numpy().view(np.uint64))
            check = bool(np.array_equal(r.cpu().numpy().view(np.uint64), want))
        except Exception as e:   # libgmp missing: leave unchecked
            check = "unchecked: %s" % e

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    L.mpirfft_launch_count_reset()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # evict the slabs from L2 between timed iterations
        evs[k][0].record()
        step()
        evs[k][1].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = int(L.mpirfft_launch_count())
    if world > 1:
        dist.barrier()
    clocks = sampler.finish()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = sum(step_ms)
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * (n1 + n2) / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the drop-in C symbol with host buffers (pinned), copies inside the timed region ----
    ha = torch.from_numpy(splitmix64(0x5EED0001 + 1000 * rank, n1).view(np.int64)).pin_memory()
    hb = torch.from_numpy(splitmix64(0x5EED0002 + 1000 * rank, n2).view(np.int64)).pin_memory()
    hr = torch.zeros(n1 + n2, dtype=torch.int64).pin_memory()
    f = L.new_mpn_mul
    for _ in range(2):
        f(hr.data_ptr(), ha.data_ptr(), n1, hb.data_ptr(), n2, depth, w)
    e2e_steps = max(3, min(args.steps, 20))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        f(hr.data_ptr(), ha.data_ptr(), n1, hb.data_ptr(), n2, depth, w)     # synchronises on return
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = world * e2e_steps * (n1 + n2) / e2e_dt / 1e6

    # ---- profiling leg: per-kernel-class CUDA events (not part of the numbers above) ----
    roofline, phases = None, None
    if rank == 0:
        nclass = 6
        ms = (C.c_double * nclass)(); ln = (C.c_uint64 * nclass)(); by = (C.c_double * nclass)()
        L.mpirfft_profile_enable(1)
        reps = 5
        for _ in range(reps):
            flush.zero_()
            step()
        L.mpirfft_profile_read(ms, ln, by, nclass)
        L.mpirfft_profile_enable(0)
        names = ["stage", "finalize", "pointwise", "split", "combine", "normalise"]
        phases = {names[i]: {"ms_per_product": ms[i] / reps, "launches_per_product": int(ln[i]) // reps} for i in range(nclass)}
        peaks, how = load_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        if ln[0]:
            avg_ms = ms[0] / ln[0]
            achieved = (by[0] / ln[0]) / (avg_ms * 1e-3) / 1e9
            roofline = {"kernel": "k_run_stage (one radix-2 layer of a transform)", "bound": "hbm",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": None, "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs)",
                        "alg_bytes_per_launch": by[0] / ln[0], "avg_launch_us": avg_ms * 1e3,
                        "share_of_step": (ms[0] / reps) / ms_per_step}
        # whole-product algorithmic traffic: 18 T S + 16 (n1+n2) bytes
        T = prm["trunc"]
        phases["B_alg_bytes"] = 18 * T * S + 16 * (n1 + n2)
        phases["B_alg_over_time_GBs"] = phases["B_alg_bytes"] / (ms_per_step * 1e-3) / 1e9
        phases["pointwise_mad32_per_product"] = 4 * prm["limbs"] ** 2 * T

    # ---- CPU baseline: the compiled reference on one host core (bounded sample) ----
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            reps = 3 if args.workload != "cfg1" else 20
            times, kind = cpu_reference_product(n1, n2, depth, w, reps)
            best = min(times)
            cpu = {"value": (n1 + n2) / best / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": "%d products of the same workload, best of %d, 1 thread (ms each: %s)" % (
                       reps, reps, ", ".join("%.0f" % (x * 1e3) for x in times)),
                   "cpu_model": _cpu_model(), "host_cores": os.cpu_count()}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": str(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "new_mpn_mul %d x %d limbs, depth %d, w %d (BASELINE configs[%s])" % (
                           n1, n2, depth, w, {"cfg1": 0, "cfg2": 1, "cfg3": 2}[args.workload]),
                       "coefficients": prm["trunc"], "limbs_per_coefficient": prm["limbs"],
                       "l2": "flushed between timed steps (256 MiB write)",
                       "multi_gpu": "independent products per rank" if world > 1 else "single GPU"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * (n1 + n2),
                    "d2h_bytes_per_step": 8 * (n1 + n2), "ms_per_step": e2e_dt / e2e_steps * 1e3,
                    "api": "new_mpn_mul(r, i1, n1, i2, n2, depth, w) with pinned host buffers"},
            "roofline": roofline, "cpu_baseline": cpu, "phases": phases,
            "bit_exact_vs_gmp": check, "wall_s_timed_region": wall,
            "step_ms_min_max": [min(step_ms), max(step_ms)],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


if __name__ == "__main__":
    main()
