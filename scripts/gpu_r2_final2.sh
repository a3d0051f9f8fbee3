#!/bin/bash
# last GPU visit of round 2: the whole GPU suite, smoke, the default bench line and cfg1 of the committed tree,
# and the big-ring drop-in call from pageable memory with and without the staging threads
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -13 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-sharded-leg > gpurun_out/bench_cfg1.log 2>&1; echo "cfg1 rc=$?"
python - <<PY
import json
for f in ("bench", "bench_cfg1"):
    l=[x for x in open("gpurun_out/%s.log" % f) if x.startswith("{")]
    d=json.loads(l[-1]); print(f, d.get("ms_per_step"), d.get("step_ms_median"), d.get("step_ms_min_max"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e_pageable") or {}).get("ms_per_step"), d.get("bit_exact_vs_gmp"), (d.get("roofline") or {}).get("frac"))
PY
for t in 6 0; do
MPIRFFT_COPY_THREADS=$t timeout 300 python - <<PY
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench, mpir_fft_b200 as M
M.init(0)
n = 1 << 24
a, b = bench.splitmix64(1, n), bench.splitmix64(2, n)
r = M.new_mpn_mul(a, b, 16, 1)
ts = []
for _ in range(3):
    t = time.perf_counter(); r = M.new_mpn_mul(a, b, 16, 1); ts.append((time.perf_counter() - t) * 1e3)
import torch
from mpir_fft_b200 import residues as R
ta, tb, tr = (torch.from_numpy(x.view(np.int64)).cuda() for x in (a, b, r))
print("COPY_THREADS=$t  16M x 16M limbs (depth 16, w 1, ring of 1024 limbs) from numpy arrays: ms", ["%.1f" % x for x in ts], "residues ok:", R.product_matches(R.residues(ta), R.residues(tb), R.residues(tr)))
PY
done
