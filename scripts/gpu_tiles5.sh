#!/bin/bash
mkdir -p gpurun_out
for c in 4 5 4 5; do
MPIRFFT_TILES_CTAS=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ctas$c.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_ctas$c.log") if x.startswith("{")]
d=json.loads(l[-1]); print("ctas $c", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "stage", d["phases"]["stage"]["ms_per_product"], d["roofline"]["frac"], d["bit_exact_vs_gmp"])
PY
done
MPIRFFT_TILES_CTAS=5 timeout 600 python -m pytest tests -m gpu -x -q -k "cfg or truncated or transform or mfa" > gpurun_out/pytest_ctas5.log 2>&1; tail -2 gpurun_out/pytest_ctas5.log
