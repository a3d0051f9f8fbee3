#!/bin/bash
# A/B of a tile-kernel switch on the default workload: $1 = environment variable, values 0 and 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in 0 1; do
for wl in cfg2 cfg1; do
env $1=$v timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-sharded-leg --no-cpu-baseline > gpurun_out/bench_${wl}_$v.log 2> gpurun_out/bench_${wl}_$v.err; echo "$1=$v bench $wl rc=$?"
grep '^{' gpurun_out/bench_${wl}_$v.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
print('ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'exact',d['bit_exact_vs_gmp'],'roofline frac',r.get('frac'),'avg_launch_us',r.get('avg_launch_us'))
print({k:(v['ms_per_product'],v['launches_per_product']) for k,v in d['phases'].items() if isinstance(v,dict) and v['launches_per_product']})
"
tail -2 gpurun_out/bench_${wl}_$v.err
done; done
