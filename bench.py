#!/usr/bin/env python
"""bench.py -- new_mpn_mul throughput on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload cfg2|cfg1|cfg3|cfg4|big] [--sharded]

A "step" is one product of the workload (default BASELINE.json configs[1]: 2^20 x 2^20 limbs,
depth 14, w 1) on synthetic uniform limbs (splitmix64 counter generator).

  value      whole-job output limbs / s, operands and result resident in HBM, CUDA-event time per
             step (L2 flushed between steps), max over ranks
  e2e        the same metric through the reference-facing C symbol with HOST buffers (pinned),
             H2D/D2H inside the timed region
  roofline   the fused transform-pass kernel (k_run_tiles, one MFA pass per launch): algorithmic
             bytes per launch / mean launch duration, from per-launch CUDA events in a separate
             profiling leg of the same run; roofline_pointwise: the product kernel against the
             IMAD issue rate measured in the same run (mpirfft_measure_imad_rate)
  cpu_baseline  the compiled reference (oracle/_ref, mul_fft.c:3246 fixed) on one host core

--impl reference times only the reference's CPU implementation (all host cores, one product per
process per step).  N > 1: one process per GPU under torchrun; each rank multiplies its own
operands (independent products, weak scaling, no data-path collective).  --sharded instead spreads
ONE product over the ranks (MFA column/row partition, NCCL all-to-all; mpir_fft_b200/sharded.py;
strong scaling) -- meant for the large workloads (`big`: 8e6 x 8e6 limbs).
--workload cfg4: 4096 independent 16 384-limb products mod 2^(2^20)+1 (BASELINE configs[3]).
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
    "big": (8000000, 8000000, 15, 1),          # l = 512: the largest ring of the fused path
}
METRIC, UNIT = "new_mpn_mul Mlimb/s", "Mlimb/s"
SHARDED_LEG_TIMEOUT_S = 420


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
        self.ready, self.marking, self.error = False, True, None

    NVML_REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                    ("sw_power_cap", 0x4))

    def _run_nvml(self):
        """in-process NVML polling (about 1 ms per sample): the timed region of one cfg2 run is tens
        of milliseconds, shorter than nvidia-smi's start-up"""
        if os.environ.get("MPIRFFT_BENCH_NO_NVML"):        # test aid: exercise the nvidia-smi fallback
            raise RuntimeError("NVML disabled by MPIRFFT_BENCH_NO_NVML")
        import pynvml as N
        N.nvmlInit()
        gpu = self.gpu
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis:
            ids = [v.strip() for v in vis.split(",")]
            if gpu < len(ids) and ids[gpu].isdigit():
                gpu = int(ids[gpu])
        h = N.nvmlDeviceGetHandleByIndex(gpu)
        mx = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        self.ready = True
        while not self.stop_flag:
            try:
                sm = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                try:
                    bits = int(N.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    bits = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                row = [str(gpu), sm, mx, "", hex(bits)] + ["Active" if bits & m else "Not Active" for _, m in self.NVML_REASONS]
                if self.marking:
                    self.rows.append([str(x) for x in row])
            except Exception as e:         # one failed query must not end the sampling
                self.error = repr(e)
            time.sleep(0.001)      # about ten samples per 20 ms of timed region; NVML queries take a driver lock

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception as e:
            self.error = repr(e)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.ready = True
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            self.ready = True

    def begin(self):
        """wait for the first sample source to be up; samples count from here"""
        t = time.time()
        while not self.ready and time.time() - t < 5.0:
            time.sleep(0.01)
        if self.proc is None:
            self.rows = []
        else:
            time.sleep(0.2)

    def finish(self):
        self.stop_flag = True
        if self.proc:
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
            # nothing arrived during the timed region (NVML unavailable, nvidia-smi too slow to start): one query
            # right behind it, while the clocks have not yet dropped
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
                r = [x.strip() for x in out[0].split(",")]
                reasons = sorted(name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9])
                                 if v.lower().startswith("active"))
                return {"sm_mhz": float(r[1]), "sm_max_mhz": float(r[2]), "reasons": reasons, "samples": 1,
                        "note": "one nvidia-smi query right after the timed region (no sample arrived during it)", "sampler_error": self.error}
            except Exception as e:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "sampler_error": "%s; %r" % (self.error, e)}
        busy = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_cpus(gpu):
    """pin this process to the CPUs next to its GPU (NVML's ideal affinity): with one process per GPU the
    host side of the end-to-end path -- staging copies, pinned buffers (first touch), launch threads --
    otherwise piles up on one NUMA node (round 1: e2e weak-scaling efficiency 0.67 at 8 GPUs)"""
    try:
        import pynvml as N
        N.nvmlInit()
        idx = gpu
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis:
            ids = [v.strip() for v in vis.split(",")]
            if gpu < len(ids) and ids[gpu].isdigit():
                idx = int(ids[gpu])
        h = N.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = N.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1 and 64 * i + b < ncpu}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "first": min(cpus), "last": max(cpus)}
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}
    return None


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
    if args.workload not in WORKLOADS:
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm times new_mpn_mul workloads only; "
                          "--workload %s reports its own cpu_baseline" % args.workload}))
        return
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
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg4"])
    ap.add_argument("--sharded", action="store_true", help="one product spread over the ranks (strong scaling)")
    ap.add_argument("--count", type=int, default=4096, help="cfg4: products per batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded-leg", action="store_true", help="skip the large sharded product that the default workload appends")
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
    numa = bind_to_gpu_cpus(local_rank)       # before any host buffer is allocated (first touch)
    M.init(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload == "cfg4":
        return run_cfg4(args, rank, local_rank, world, torch, dist, M)
    if args.sharded:
        return run_sharded(args, rank, local_rank, world, torch, dist, M)
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
    sampler.begin()
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

    # ---- the library's own parameter choice (mpirfft_mpn_mul: may pick the sqrt2 shape of new_mpn_mul6) ----
    chooser = None
    try:
        cd, cw, csq = M.choose_params6(n1, n2)
        if (cd, cw, csq) != (depth, w, False):
            plan2 = M.MulPlan(n1, n2, cd, cw, sqrt2=csq)
            r2 = torch.zeros(n1 + n2, dtype=torch.int64, device="cuda")
            for _ in range(args.warmup):
                plan2.exec_device(r2.data_ptr(), pa, pb, None)
            torch.cuda.synchronize()
            ok2 = bool(np.array_equal(r2.cpu().numpy().view(np.uint64), want)) if (rank == 0 and check is True) else None
            ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for k in range(args.steps):
                flush.zero_()
                ev2[k][0].record()
                plan2.exec_device(r2.data_ptr(), pa, pb, None)
                ev2[k][1].record()
            torch.cuda.synchronize()
            ms2 = sum(a_.elapsed_time(b_) for a_, b_ in ev2) / args.steps
            chooser = {"api": "mpirfft_mpn_mul / mpirfft_choose_params6", "depth": cd, "w": cw, "sqrt2_new_mpn_mul6": csq,
                       "limbs_per_coefficient": plan2.params["limbs"], "coefficients": plan2.params["trunc"],
                       "ms_per_step": ms2, "value": (n1 + n2) / (ms2 * 1e-3) / 1e6, "unit": UNIT, "bit_exact_vs_gmp": ok2,
                       "launches_per_product": plan2.launches}
            plan2.close()
            del r2
        else:
            chooser = {"api": "mpirfft_mpn_mul / mpirfft_choose_params6", "depth": cd, "w": cw, "sqrt2_new_mpn_mul6": csq,
                       "note": "same parameters as the workload"}
    except Exception as e:
        chooser = {"error": "%s: %s" % (type(e).__name__, e)}

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
    e2e_check = None
    if rank == 0 and check is True:      # what the timed host-buffer calls returned, against the same GMP product
        e2e_check = bool(np.array_equal(hr.numpy().view(np.uint64), want))
    # the same symbol on plain malloc'ed (pageable) buffers -- what a caller that swaps libraries hands in
    na, nb = splitmix64(0x5EED0001 + 1000 * rank, n1), splitmix64(0x5EED0002 + 1000 * rank, n2)
    nr = np.zeros(n1 + n2, dtype=np.uint64)
    for _ in range(2):
        f(nr.ctypes.data, na.ctypes.data, n1, nb.ctypes.data, n2, depth, w)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        f(nr.ctypes.data, na.ctypes.data, n1, nb.ctypes.data, n2, depth, w)
    e2e_pg_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_pg_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_pg_dt = float(t.item())
    e2e_pg_value = world * e2e_steps * (n1 + n2) / e2e_pg_dt / 1e6
    e2e_pg_check = bool(np.array_equal(nr, want)) if (rank == 0 and check is True) else None

    # ---- profiling leg: per-kernel-class CUDA events (not part of the numbers above) ----
    roofline, phases, roofline_pw = None, None, None
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
        traffic = None
        try:    # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (cfg2)
            if args.workload == "cfg2":
                import csv
                with open(os.path.join(ROOT, "profiles", "r01_ncu_tiles_summary.csv")) as f:
                    rows = list(csv.reader(f))
                ir, iw = rows[0].index("dram__bytes_read.sum"), rows[0].index("dram__bytes_write.sum")
                vals = [(float(r[ir]) + float(r[iw])) * 1e6 for r in rows[2:] if len(r) > iw]
                traffic = sum(vals) / len(vals) if vals else None
        except Exception:
            traffic = None
        if ln[0]:
            avg_ms = ms[0] / ln[0]
            achieved = (by[0] / ln[0]) / (avg_ms * 1e-3) / 1e9
            roofline = {"kernel": "k_run_tiles (one fused MFA pass: several radix-2 layers in shared memory)", "bound": "hbm",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "traffic_source": "profiles/r02_ncu_tiles_summary.csv: mean dram read+write bytes "
                        "per launch over the captured passes (ncu --set full, cold L2)" if traffic else None,
                        "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs)",
                        "alg_bytes_per_launch": by[0] / ln[0], "avg_launch_us": avg_ms * 1e3,
                        "share_of_step": (ms[0] / reps) / ms_per_step}
        # whole-product algorithmic traffic: 18 T S + 16 (n1+n2) bytes
        T = prm["trunc"]
        phases["B_alg_bytes"] = 18 * T * S + 16 * (n1 + n2)
        phases["B_alg_over_time_GBs"] = phases["B_alg_bytes"] / (ms_per_step * 1e-3) / 1e9
        phases["pointwise_mad32_per_product"] = 4 * prm["limbs"] ** 2 * T
        # the product kernel against the integer-multiply issue rate measured right here
        try:
            L.mpirfft_measure_imad_rate.restype = C.c_double
            imad_chain = float(L.mpirfft_measure_imad_rate(1))      # IMAD.WIDE.U32.X carry chains (what the kernel issues)
            imad_wide = float(L.mpirfft_measure_imad_rate(0))       # IMAD.WIDE.U32(.., RZ) + IADD3 + IADD3.X: the split form ptxas emits for mad.wide.u32
            pw_ms = phases["pointwise"]["ms_per_product"]
            if pw_ms > 0 and imad_chain > 0:
                school = phases["pointwise_mad32_per_product"] / (pw_ms * 1e-3)
                # l = 128 / 256 / 512: every block product is split once (Karatsuba): 3/4 of the multiply-adds are issued
                kara = prm["limbs"] in (128, 256, 512) and os.environ.get("MPIRFFT_POINTWISE", "") in ("", "k")
                issued = school * (0.75 if kara else 1.0)
                roofline_pw = {"kernel": "k_pointwise (%s block products, 32x32->64 multiply-add carry chains)" % (
                                   "Karatsuba-split" if kara else "schoolbook"), "bound": "imad",
                               "achieved": issued / 1e12, "peak": imad_chain / 1e12, "unit": "Tmad32/s",
                               "frac": issued / imad_chain,
                               "schoolbook_equivalent": school / 1e12, "schoolbook_equivalent_frac": school / imad_chain,
                               "imad_wide_plus_iadd3_rate": imad_wide / 1e12,
                               "note": "achieved = multiply-adds actually issued (utilisation); schoolbook_equivalent = 4 l^2 per "
                                       "product / time (effective).  peak = IMAD.WIDE.U32.X carry chains, the only fused "
                                       "32x32+64 form ptxas emits on sm_100a (mad.wide.u32 becomes IMAD.WIDE(..,RZ) + IADD3 + IADD3.X)",
                               "peak_source": "measured in this run (csrc/cuda/imad_peak.cu)"}
        except Exception:
            roofline_pw = None

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

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "new_mpn_mul %d x %d limbs, depth %d, w %d" % (n1, n2, depth, w),
                       "baseline_config_index": {"cfg1": 0, "cfg2": 1, "cfg3": 2, "big": None}[args.workload],
                       "coefficients": prm["trunc"], "limbs_per_coefficient": prm["limbs"],
                       "l2": "flushed between timed steps (256 MiB write)",
                       "multi_gpu": "independent products per rank" if world > 1 else "single GPU",
                       "host_affinity": numa},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * (n1 + n2),
                    "d2h_bytes_per_step": 8 * (n1 + n2), "ms_per_step": e2e_dt / e2e_steps * 1e3,
                    "api": "new_mpn_mul(r, i1, n1, i2, n2, depth, w) with pinned host buffers",
                    "bit_exact_vs_gmp": e2e_check},
            "e2e_pageable": {"value": e2e_pg_value, "unit": UNIT, "ms_per_step": e2e_pg_dt / e2e_steps * 1e3,
                             "api": "new_mpn_mul(r, i1, n1, i2, n2, depth, w) with malloc'ed (pageable) host buffers",
                             "bit_exact_vs_gmp": e2e_pg_check},
            "roofline": roofline, "roofline_pointwise": roofline_pw, "cpu_baseline": cpu, "phases": phases,
            "library_parameter_choice": chooser,
            "bit_exact_vs_gmp": check, "wall_s_timed_region": wall,
            "step_ms_min_max": [min(step_ms), max(step_ms)],
            "step_ms_median": sorted(step_ms)[len(step_ms) // 2], "slowest_step_index": step_ms.index(max(step_ms)),
        }

    # ---- the collective path: ONE large product sharded over all ranks (strong scaling) ----
    # Appended to the same line as "sharded".  A watchdog prints the line without it if a rank gets
    # stuck, so the headline numbers above can never be lost to the large leg.
    if args.workload == "cfg2" and not args.no_sharded_leg:
        def bail():
            if rank == 0:
                line["sharded"] = {"error": "the sharded leg did not finish within %d s" % SHARDED_LEG_TIMEOUT_S}
                print(json.dumps(line), flush=True)
            os._exit(0)
        dog = threading.Timer(SHARDED_LEG_TIMEOUT_S, bail)
        dog.daemon = True
        dog.start()
        recs = []
        try:
            del a, b, r, flush
            plan.close() if hasattr(plan, "close") else None
            torch.cuda.empty_cache()
            peaks, _ = load_peaks()
            for lg in SHARDED_SIZES.get(world, (26,)):
                rec = sharded_leg(torch, dist, M, rank, world, lg, steps=2, peak_gbs=float(peaks.get("hbm_gbs", 6650.0)))
                if rank == 0:
                    recs.append(rec)
            if rank == 0 and world == 1 and not args.no_cpu_baseline and recs:
                # the reference on one host core at a bounded sample of the same kind of workload
                try:
                    times, kind = cpu_reference_product(1 << 24, 1 << 24, 16, 1, 1)
                    recs[0]["cpu_baseline"] = {"value": (2 << 24) / times[0] / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
                                               "sample": "one 2^24 x 2^24-limb product (depth 16, w 1; a quarter of the operand size), "
                                                         "%.1f s, 1 thread" % times[0], "cpu_model": _cpu_model(), "host_cores": os.cpu_count()}
                except Exception as e:
                    recs[0]["cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": str(e)}
        except Exception as e:      # peers may now be waiting in a collective: the watchdog ends the run
            import traceback
            recs.append({"error": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-600:]})
            if world > 1:
                if rank == 0:
                    line["sharded"] = recs
                    print(json.dumps(line), flush=True)
                os._exit(0)
        dog.cancel()
        if rank == 0:
            line["sharded"] = recs
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _timed_steps(torch, dist, world, steps, flush, fn):
    """K steps, each bracketed by CUDA events on the current stream, L2 flushed in between"""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.zero_()
        evs[k][0].record()
        fn()
        evs[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = [a.elapsed_time(b) for a, b in evs]
    total = sum(ms)
    if world > 1:
        t = torch.tensor([total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())
    return total / steps, ms


# one large product spread over the ranks (SURVEY 8e / BASELINE configs[4]); log2 limbs -> (depth, w)
SHARDED_PARAMS = {24: (16, 1), 26: (17, 1), 28: (18, 1), 30: (19, 1)}
# which sizes a run with `world` ranks measures: 2^26 everywhere (1-GPU baseline on rank 0 for the
# strong-scaling efficiency), plus the largest size the box holds
SHARDED_SIZES = {1: (26,), 2: (26,), 4: (26, 28), 8: (26, 30)}


def splitmix64_dev(torch, seed, n, dev, chunk=1 << 26):
    """limb[k] = splitmix64(seed + k) as int64 bit patterns, generated on the device in chunks"""
    out = torch.empty(n, dtype=torch.int64, device=dev)

    def lsr(z, sh):
        return (z >> sh) & ((1 << (64 - sh)) - 1)

    def c64(v):
        return v - (1 << 64) if v >= (1 << 63) else v

    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        z = torch.arange(lo, hi, dtype=torch.int64, device=dev) + c64(seed & ((1 << 64) - 1))
        z = z * c64(0x9E3779B97F4A7C15) + c64(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * c64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * c64(0x94D049BB133111EB)
        out[lo:hi] = z ^ lsr(z, 31)
    return out


def _phase_table(torch, timing):
    """[(name, event)] -> {name: ms} summed over repeated phases (device time on the current stream)"""
    out = {}
    for (_, e0), (name, e1) in zip(timing[:-1], timing[1:]):
        out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
    return out


def sharded_leg(torch, dist, M, rank, world, log2, steps=2, baseline_1gpu=True, peak_gbs=None):
    """ONE product of 2^log2 x 2^log2 limbs over all ranks: time, phases, all-to-all rate, bit-exactness
    through residues modulo six 31-bit primes (mpir_fft_b200/residues.py) and, at 2^26, equality of the
    sharded result's fingerprint with the same product computed on rank 0 alone."""
    from mpir_fft_b200.sharded import ShardedMul
    from mpir_fft_b200 import residues as RES
    depth, w = SHARDED_PARAMS[log2]
    n = 1 << log2
    dev = torch.device("cuda", torch.cuda.current_device())
    a = splitmix64_dev(torch, 0x5EED0001, n, dev)
    b = splitmix64_dev(torch, 0x5EED0002, n, dev)
    # residues of the operands: every rank reduces its slice
    lo, hi = rank * n // world, (rank + 1) * n // world
    ra, rb = RES.residues(a[lo:hi], lo), RES.residues(b[lo:hi], lo)

    def allsum(vals):
        if world == 1:
            return [v % p for v, p in zip(vals, RES.PRIMES)]
        t = torch.tensor(vals, dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        return [int(v) % p for v, p in zip(t.cpu().tolist(), RES.PRIMES)]

    ra, rb = allsum(ra), allsum(rb)
    sm = ShardedMul(n, n, depth, w, cuda=True)
    lay = sm.lay
    sm.multiply(a.data_ptr(), b.data_ptr())             # warm-up (NCCL channels, plan buffers)
    torch.cuda.synchronize()
    times = []
    for _ in range(steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sm.multiply(a.data_ptr(), b.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    rr = allsum(RES.residues(sm.out, int(lay.limb_lo)))
    ok = RES.product_matches(ra, rb, rr)
    # one more product with a CUDA event after every phase
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    timing = []
    sm.multiply(a.data_ptr(), b.data_ptr(), timing=timing)
    torch.cuda.synchronize()
    phases = _phase_table(torch, timing)
    pt = torch.tensor([phases.get(k, 0.0) for k in sorted(phases)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
    phases = {k: float(v) for k, v in zip(sorted(phases), pt.cpu().tolist())}
    # end to end with HOST operands (pinned): every rank copies the operands in, the product runs, the
    # rank's window of the result is copied out -- all inside the timed region (sizes that fit host memory)
    e2e = None
    if log2 <= 26:
        ha = torch.empty(n, dtype=torch.int64).pin_memory(); ha.copy_(a)
        hb = torch.empty(n, dtype=torch.int64).pin_memory(); hb.copy_(b)
        hout = torch.empty(sm.out.numel(), dtype=torch.int64).pin_memory()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.copy_(ha, non_blocking=True)
        b.copy_(hb, non_blocking=True)
        sm.multiply(a.data_ptr(), b.data_ptr())
        hout.copy_(sm.out, non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        e2e = {"value": 2 * n / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": 16 * n * world, "d2h_bytes_per_step": 16 * n,
               "api": "ShardedMul.multiply (mpirfft_smul_* phases + NCCL) with pinned host operands copied to every rank "
                      "and the result windows copied back"}
        del ha, hb, hout
    a2a_bytes = int(lay.trunc_rows * lay.ncl * lay.block_limbs * 8)          # one rank's buffer per exchange
    sent = a2a_bytes * (world - 1) // world                                  # what actually leaves the GPU
    mem = torch.cuda.max_memory_allocated() / 1e9
    sm.close()
    del sm
    base = None
    if baseline_1gpu and world > 1 and log2 <= 26:      # (the larger sizes do not fit one GPU's plan buffers)
        # the same product on rank 0 alone: strong-scaling denominator and world-size independence
        same = None
        if rank == 0:
            s1 = ShardedMul(n, n, depth, w, cuda=True, single=True)
            s1.multiply(a.data_ptr(), b.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.multiply(a.data_ptr(), b.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            base = e0.elapsed_time(e1)
            same = (RES.residues(s1.out, 0) == rr)
            s1.close()
            del s1
        dist.barrier()
    else:
        same = None
    del a, b
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    ms = min(times)
    S = int(lay.block_limbs - 1) * 8        # pitch = l + 2 limbs; S = 8 (l + 1)
    T = int(lay.trunc_rows) * int(lay.n1cols)
    b_alg = 18 * T * S + 16 * 2 * n
    a2a_ms = phases.get("all_to_all", 0.0)
    rec = {
        "workload": "new_mpn_mul 2^%d x 2^%d limbs, depth %d, w %d, ONE product sharded over %d rank(s)" % (log2, log2, depth, w, world),
        "n_gpus": world, "ms_per_step": ms, "all_steps_ms": times, "value": 2 * n / (ms * 1e-3) / 1e6, "unit": UNIT,
        "coefficient_limbs": int(lay.block_limbs) - 2, "coefficients": T,
        "phases_ms": phases,
        "all_to_all": {"exchanges_per_product": 3, "buffer_bytes_per_rank_per_exchange": a2a_bytes,
                       "sent_bytes_per_rank_per_exchange": sent, "ms_total": a2a_ms,
                       "GBs_per_rank_unidirectional": (3 * sent / (a2a_ms * 1e-3) / 1e9) if (a2a_ms > 0 and world > 1) else None},
        "ms_1gpu_same_product": base,
        "strong_efficiency_vs_1gpu": (base / (world * ms)) if base else None,
        "roofline": {"bound": "hbm", "what": "whole product: SURVEY 8(d) algorithmic bytes 18 T S + 16 (n1+n2) over all ranks",
                     "alg_bytes": b_alg, "achieved": b_alg / (ms * 1e-3) / 1e9, "unit": "GB/s",
                     "peak": (peak_gbs * world) if peak_gbs else None,
                     "frac": (b_alg / (ms * 1e-3) / 1e9 / (peak_gbs * world)) if peak_gbs else None},
        "e2e": e2e,
        "bit_exact": bool(ok and (same is not False)),
        "check": {"residues": "a*b == r modulo the six 31-bit primes of mpir_fft_b200/residues.py (weights 2^(64k) mod p, "
                              "order of 2 > 3e8: position-sensitive), result limbs reduced where they live",
                  "residues_match": bool(ok), "result_fingerprint": rr,
                  "equals_1gpu_result": same,
                  "full_compare": "2^26 x 2^26 against GMP mpn_mul limb for limb: profiles/r02_full_compare_2p26.json (scripts/full_compare.py)"},
        "device_mem_gb": mem,
    }
    return rec


def run_cfg4(args, rank, local_rank, world, torch, dist, M):
    """BASELINE configs[3]: `count` independent 16 384-limb products mod 2^(2^20)+1 per step and rank
    (weak scaling over ranks: independent batches, no collective)"""
    Lb = M.lib()
    l, count = 16384, args.count
    vp, u64 = C.c_void_p, C.c_uint64
    Lb.mpirfft_mulmod_params.argtypes = [C.c_long, C.POINTER(u64), C.POINTER(u64)]
    Lb.mpirfft_mulmod_plan_create.argtypes = [C.POINTER(vp), C.c_long, u64, u64, C.c_size_t]
    Lb.mpirfft_mulmod_plan_exec.argtypes = [vp, vp, vp, vp, C.c_size_t, C.c_int, vp]
    Lb.mpirfft_mulmod_plan_destroy.argtypes = [vp]
    d, w = u64(), u64()
    assert Lb.mpirfft_mulmod_params(l, C.byref(d), C.byref(w)) == 0
    plan = vp()
    rc = Lb.mpirfft_mulmod_plan_create(C.byref(plan), l, d, w, count)
    if rc != 0:
        raise SystemExit("mpirfft_mulmod_plan_create failed (%d): %s" % (rc, Lb.mpirfft_last_error().decode()))
    pitch = l + 2
    ha = splitmix64(0x5EED0004 + 1000 * rank, count * pitch).reshape(count, pitch)
    hb = splitmix64(0x5EED0005 + 1000 * rank, count * pitch).reshape(count, pitch)
    ha[:, l:] = 0
    hb[:, l:] = 0
    a = torch.from_numpy(ha.view(np.int64)).cuda()
    b = torch.from_numpy(hb.view(np.int64)).cuda()
    r = torch.zeros_like(a)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        rc = Lb.mpirfft_mulmod_plan_exec(plan, r.data_ptr(), a.data_ptr(), b.data_ptr(), pitch, -1, None)
        assert rc == 0, Lb.mpirfft_last_error()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    check = None
    if rank == 0:       # a sample of the batch against GMP (mpn_mul + fold), outside the timed region
        from oracle import loader as oracle
        out = r.cpu().numpy().view(np.uint64)
        p = (1 << (64 * l)) + 1
        check = True
        for k in (0, count // 2, count - 1):
            full = oracle.gmp_mul(ha[k, :l].copy(), hb[k, :l].copy())
            want = (int.from_bytes(full[:l].tobytes(), "little") - int.from_bytes(full[l:].tobytes(), "little")) % p
            got = int.from_bytes(out[k, :l].tobytes(), "little") + (int(out[k, l]) << (64 * l))
            check = check and (got == want)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.begin()
    Lb.mpirfft_launch_count_reset()
    ms_per_step, step_ms = _timed_steps(torch, dist, world, args.steps, flush, step)
    launches = int(Lb.mpirfft_launch_count())
    clocks = sampler.finish()
    value = world * count * l / (ms_per_step * 1e-3) / 1e6
    # phases
    phase_ms = []
    for ph in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        Lb.mpirfft_mulmod_plan_exec(plan, r.data_ptr(), a.data_ptr(), b.data_ptr(), pitch, ph, None)
        e1.record()
        torch.cuda.synchronize()
        phase_ms.append(e0.elapsed_time(e1))
    # e2e: host blocks in pinned memory, copies inside the timed region
    pa, pb = torch.from_numpy(ha.view(np.int64)).pin_memory(), torch.from_numpy(hb.view(np.int64)).pin_memory()
    pr = torch.zeros_like(pa).pin_memory()
    e2e_steps = max(3, min(args.steps, 5))
    t0 = None
    for it in range(e2e_steps + 1):
        if it == 1:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        a.copy_(pa, non_blocking=True)
        b.copy_(pb, non_blocking=True)
        step()
        pr.copy_(r, non_blocking=True)
        torch.cuda.synchronize()
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    # CPU baseline: the compiled reference's fft_mulmod_2expp1 on one core, bounded sample
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle import loader as oracle
            ref = oracle.load_ref(True)
            P = lambda x: C.c_void_p(x.ctypes.data)   # noqa: E731
            rr, tt = np.zeros(l + 1, np.uint64), np.zeros(2 * l + 2, np.uint64)
            nsample, times = 200, []
            for k in range(nsample):
                x, y = ha[k % count, :l + 1].copy(), hb[k % count, :l + 1].copy()
                t = time.perf_counter()
                ref.fft_mulmod_2expp1(P(rr), P(x), P(y), C.c_long(1 << 14), C.c_long(64), P(tt))
                times.append(time.perf_counter() - t)
            cpu = {"value": l / statistics.median(times) / 1e6, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": "%d of the %d products, median per product %.2f ms, 1 thread" % (nsample, count, statistics.median(times) * 1e3),
                   "cpu_model": _cpu_model(), "host_cores": os.cpu_count()}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": str(e)}
    if rank == 0:
        inner_l = ((1 << d.value) * w.value) // 64
        mads = count * (2 << d.value) * 4 * inner_l ** 2
        try:    # the product kernel against the integer multiply-add issue rate measured right here
            Lb.mpirfft_measure_imad_rate.restype = C.c_double
            imad_chain = float(Lb.mpirfft_measure_imad_rate(1))
        except Exception:
            imad_chain = 0.0
        kara = inner_l in (128, 256, 512)
        school = mads / (phase_ms[2] * 1e-3)
        issued = school * (0.75 if kara else 1.0)
        cfg4_roofline = {"kernel": "k_pointwise (inner products mod 2^%d+1, %s block products)" % (
                             64 * inner_l, "Karatsuba-split" if kara else "schoolbook"), "bound": "imad",
                         "achieved": issued / 1e12, "unit": "Tmad32/s", "peak": (imad_chain / 1e12) if imad_chain > 0 else None,
                         "frac": (issued / imad_chain) if imad_chain > 0 else None,
                         "schoolbook_equivalent": school / 1e12,
                         "schoolbook_equivalent_frac": (school / imad_chain) if imad_chain > 0 else None,
                         "traffic": None, "share_of_step": phase_ms[2] / ms_per_step,
                         "peak_source": "IMAD.WIDE.U32.X carry-chain rate measured in this run (csrc/cuda/imad_peak.cu)"}
        print(json.dumps({
            "metric": "fft_mulmod_2expp1 batch Mlimb/s", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "%d independent products of 16384 limbs mod 2^(2^20)+1 (BASELINE configs[3]); inner ring "
                                   "2^%d+1, %d pieces" % (count, (1 << d.value) * w.value, 2 << d.value),
                       "l2": "flushed between timed steps (256 MiB write); batch is 1.1 GB per operand",
                       "multi_gpu": "independent batches per rank" if world > 1 else "single GPU"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": world * count * l / e2e_dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": 2 * ha.nbytes,
                    "d2h_bytes_per_step": ha.nbytes, "ms_per_step": e2e_dt * 1e3,
                    "api": "mpirfft_mulmod_plan_exec on HBM blocks, pinned host blocks copied in and out every step"},
            "roofline": cfg4_roofline,
            "cpu_baseline": cpu,
            "phases": {"split+forward a": phase_ms[0], "split+forward b": phase_ms[1], "pointwise": phase_ms[2],
                       "inverse": phase_ms[3], "finish": phase_ms[4], "unit": "ms per batch"},
            "bit_exact_vs_gmp": check, "step_ms_min_max": [min(step_ms), max(step_ms)],
        }))
    Lb.mpirfft_mulmod_plan_destroy(plan)
    if world > 1:
        dist.destroy_process_group()


def run_sharded(args, rank, local_rank, world, torch, dist, M):
    """one product spread over the ranks (strong scaling): MFA columns / rows partitioned, NCCL
    all-to-all between the passes"""
    from mpir_fft_b200.sharded import ShardedMul
    n1, n2, depth, w = WORKLOADS[args.workload]
    a = torch.from_numpy(splitmix64(0x5EED0001, n1).view(np.int64)).cuda()
    b = torch.from_numpy(splitmix64(0x5EED0002, n2).view(np.int64)).cuda()
    sm = ShardedMul(n1, n2, depth, w, cuda=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        sm.multiply(a.data_ptr(), b.data_ptr())

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    res = sm.gather_result()
    check = None
    if rank == 0:
        from oracle import loader as oracle
        check = bool(np.array_equal(res, oracle.gmp_mul(a.cpu().numpy().view(np.uint64), b.cpu().numpy().view(np.uint64))))
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.begin()
    M.lib().mpirfft_launch_count_reset()
    ms_per_step, step_ms = _timed_steps(torch, dist, world, args.steps, flush, step)
    launches = int(M.lib().mpirfft_launch_count())
    clocks = sampler.finish()
    if rank == 0:
        lay = sm.lay
        print(json.dumps({
            "metric": METRIC, "value": (n1 + n2) / (ms_per_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "new_mpn_mul %d x %d limbs, depth %d, w %d, ONE product sharded over %d rank(s)" % (n1, n2, depth, w, world),
                       "l2": "flushed between timed steps (256 MiB write)",
                       "multi_gpu": "MFA columns/rows partitioned, 3 all-to-alls + halo all-gather + carry hand-off per product",
                       "all_to_all_bytes_per_rank": int(lay.trunc_rows * lay.ncl * lay.block_limbs * 8)},
            "clocks": clocks, "gpu_launches": launches, "e2e": None, "roofline": None, "cpu_baseline": None,
            "bit_exact_vs_gmp": check, "step_ms_min_max": [min(step_ms), max(step_ms)],
        }))
    sm.close()
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
