#!/bin/bash
# quick GPU check of a kernel change: parity tests, then the default bench line without the large sharded leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for wl in cfg2 cfg1 cfg3; do
timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-sharded-leg --no-cpu-baseline > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
grep '^{' gpurun_out/bench_$wl.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
print('ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e_pageable']['ms_per_step'],'exact',d['bit_exact_vs_gmp'],d['e2e']['bit_exact_vs_gmp'])
print('roofline frac',r.get('frac'),'avg_launch_us',r.get('avg_launch_us'),'launches',d['gpu_launches'])
print({k:(v['ms_per_product'],v['launches_per_product']) for k,v in d['phases'].items() if isinstance(v,dict)})
"
tail -2 gpurun_out/bench_$wl.err
done
