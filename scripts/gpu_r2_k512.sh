#!/bin/bash
# A/B at 512-limb coefficients: schoolbook block products (d) vs two-sweep Karatsuba (k), step-loop unroll 1 / 4
mkdir -p gpurun_out
MPIRFFT_POINTWISE=k timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mulmod or cfg3 or odd_sizes" > gpurun_out/pytest_k512.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_k512.log
for v in d k; do for u in 1 4; do
for wl in cfg3 big; do
env MPIRFFT_POINTWISE=$v MPIRFFT_PW_UNROLL=$u timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-sharded-leg --no-cpu-baseline > gpurun_out/bench_${wl}_$v$u.log 2> gpurun_out/bench_${wl}_$v$u.err; echo "POINTWISE=$v UNROLL=$u bench $wl rc=$?"
grep '^{' gpurun_out/bench_${wl}_$v$u.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'exact',d['bit_exact_vs_gmp'], 'lib', (d.get('library_parameter_choice') or {}).get('ms_per_step'))
print({k:(v['ms_per_product'],v['launches_per_product']) for k,v in d['phases'].items() if isinstance(v,dict) and v['launches_per_product']})
"
tail -2 gpurun_out/bench_${wl}_$v$u.err
done; done; done
