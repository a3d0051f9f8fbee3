#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 0 1; do
MPIRFFT_COMBINE_FUSED=$f timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused$f.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_fused$f.log") if x.startswith("{")]
d=json.loads(l[-1]); print("fused $f", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("bit_exact_vs_gmp"), "combine", d["phases"]["combine"], d["bit_exact_vs_gmp"], d["gpu_launches"])
PY
done
MPIRFFT_COMBINE_FUSED=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_fused.log 2>&1; tail -2 gpurun_out/pytest_fused.log
