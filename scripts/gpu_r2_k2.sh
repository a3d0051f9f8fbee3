#!/bin/bash
# pointwise product kernel: one Karatsuba level (k) vs two (2: one form per sweep, m: L and Hh in one sweep), unroll 1 / 4
mkdir -p gpurun_out
for m in 2 m; do
MPIRFFT_POINTWISE=$m timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mulmod or cfg or odd_sizes or mul6" > gpurun_out/pytest_k2_$m.log 2>&1; echo "pytest POINTWISE=$m rc=$?"; tail -2 gpurun_out/pytest_k2_$m.log
done
run() { # mode unroll workload
env MPIRFFT_POINTWISE=$1 MPIRFFT_PW_UNROLL=$2 timeout 600 python bench.py --workload $3 --steps 10 --warmup 3 --no-sharded-leg --no-cpu-baseline > gpurun_out/bench_$3_$1$2.log 2> gpurun_out/bench_$3_$1$2.err; echo "POINTWISE=$1 UNROLL=$2 bench $3 rc=$?"
grep '^{' gpurun_out/bench_$3_$1$2.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'exact',d['bit_exact_vs_gmp'], 'lib', (d.get('library_parameter_choice') or {}).get('ms_per_step'))
print({k:(v['ms_per_product'],v['launches_per_product']) for k,v in d['phases'].items() if isinstance(v,dict) and v['launches_per_product']})
"
tail -2 gpurun_out/bench_$3_$1$2.err
}
for m in k 2 m; do for u in 1 4; do run $m $u cfg2; done; done
for m in k 2; do for u in 1 4; do run $m $u cfg3; done; done
run 2 4 big
