#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/pointwise_modes3.log
for u in 1 4; do
MPIRFFT_PW_UNROLL=$u timeout 300 python scripts/pointwise_modes.py others >> gpurun_out/pointwise_modes3.log 2>&1; echo "modes unroll $u rc=$?"
MPIRFFT_PW_UNROLL=$u timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/cfg4_u$u.log 2>&1; echo "cfg4 unroll $u rc=$?"
python - <<PY
import json
l=[x for x in open("gpurun_out/cfg4_u$u.log") if x.startswith("{")]
d=json.loads(l[-1]); print("cfg4 unroll $u", d.get("ms_per_step"))
PY
done
cat gpurun_out/pointwise_modes3.log
for p in 0 1 0 1; do
MPIRFFT_PDL=$p timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$p.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_pdl$p.log") if x.startswith("{")]
d=json.loads(l[-1]); print("pdl $p", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["phases"]["pointwise"]["ms_per_product"], d["phases"]["combine"]["ms_per_product"], d["bit_exact_vs_gmp"])
PY
done
