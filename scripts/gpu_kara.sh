#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/pointwise_modes.py > gpurun_out/pointwise_modes.log 2>&1; echo "modes rc=$?"
cat gpurun_out/pointwise_modes.log
timeout 600 python -m pytest tests -m gpu -x -q -k "mulmod or cfg" > gpurun_out/pytest_kara.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_kara.log
for m in d k; do
MPIRFFT_POINTWISE=$m timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/cfg4_$m.log 2>&1; echo "cfg4 $m rc=$?"
python - <<PY
import json
l=[x for x in open("gpurun_out/cfg4_$m.log") if x.startswith("{")]
d=json.loads(l[-1]); print("cfg4 $m", d.get("ms_per_step"), d.get("clocks"))
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_a.log
MPIRFFT_PDL=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl.log 2> gpurun_out/bench_pdl.err; echo "bench pdl rc=$?"
python - <<PY
import json
for f in ("bench_a", "bench_pdl"):
    l=[x for x in open("gpurun_out/%s.log" % f) if x.startswith("{")]
    d=json.loads(l[-1]); print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases"]["stage"], d["roofline"]["frac"], d["bit_exact_vs_gmp"])
PY
MPIRFFT_PDL=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_pdl.log 2>&1; echo "pytest pdl rc=$?"
tail -3 gpurun_out/pytest_pdl.log
