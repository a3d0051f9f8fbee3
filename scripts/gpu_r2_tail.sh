#!/bin/bash
# product kernel: remainder of the last wave in a second launch of one-warp CTAs (default) vs one launch (MPIRFFT_PW_TAIL=0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mulmod or cfg or squaring or mul6 or wrapper" > gpurun_out/pytest_tail.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_tail.log
for v in 1 0 1 0; do
for wl in cfg2 cfg3; do
MPIRFFT_PW_TAIL=$v timeout 600 python bench.py --workload $wl --steps 30 --warmup 3 --no-sharded-leg --no-cpu-baseline > gpurun_out/bench_${wl}_tail$v.log 2> gpurun_out/bench_${wl}_tail$v.err; echo "PW_TAIL=$v bench $wl rc=$?"
grep '^{' gpurun_out/bench_${wl}_tail$v.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms',d['ms_per_step'],'median',d['step_ms_median'],'e2e',d['e2e']['ms_per_step'],'exact',d['bit_exact_vs_gmp'],'lib',(d.get('library_parameter_choice') or {}).get('ms_per_step'),'launches',d['gpu_launches'])
print({k:(v['ms_per_product'],v['launches_per_product']) for k,v in d['phases'].items() if isinstance(v,dict) and v['launches_per_product']})
"
tail -2 gpurun_out/bench_${wl}_tail$v.err
done; done
