#!/bin/bash
# round-2 GPU visit: parity tests, smoke, bench lines (all workloads + reference arm), ncu launch list,
# ncu --set full of the transform-pass and the product kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
if [ "$1" != "notests" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_reference.log
for wl in cfg1 cfg3 big; do
timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-sharded-leg > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"
done
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4.log 2>&1; echo "cfg4 rc=$?"
python - <<PY
import json
for f in ("bench", "bench_cfg1", "bench_cfg3", "bench_big", "bench_cfg4"):
    l=[x for x in open("gpurun_out/%s.log" % f) if x.startswith("{")]
    d=json.loads(l[-1]); print(f, d.get("ms_per_step"), (d.get("e2e") or {}).get("ms_per_step"), d.get("bit_exact_vs_gmp"), (d.get("roofline") or {}).get("frac"), json.dumps(d.get("library_parameter_choice"))[:300])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sharded-leg > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pointwise -s 3 -c 1 -o gpurun_out/prof_pw -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sharded-leg > gpurun_out/ncu_pw.log 2>&1; echo "ncu pw rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_tiles -s 44 -c 5 -o gpurun_out/prof_tiles -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sharded-leg > gpurun_out/ncu_full.log 2>&1; echo "ncu tiles rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
