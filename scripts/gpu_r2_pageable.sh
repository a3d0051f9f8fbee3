#!/bin/bash
# new_mpn_mul from / to pageable host memory: worker-thread staging (hostcopy.c) vs plain copies (MPIRFFT_COPY_THREADS=0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg or odd_sizes or wrapper or mul6 or adversarial or big_ring" > gpurun_out/pytest_pageable.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_pageable.log
for t in 0 2 4 6 8 12; do
for wl in cfg2 cfg1 cfg3; do
MPIRFFT_COPY_THREADS=$t timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-sharded-leg --no-cpu-baseline > gpurun_out/bench_${wl}_ct$t.log 2> gpurun_out/bench_${wl}_ct$t.err; echo "COPY_THREADS=$t bench $wl rc=$?"
grep '^{' gpurun_out/bench_${wl}_ct$t.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms',d['ms_per_step'],'e2e pinned',d['e2e']['ms_per_step'],'e2e pageable',(d.get('e2e_pageable') or {}).get('ms_per_step'),'exact',d['bit_exact_vs_gmp'],(d.get('e2e_pageable') or {}).get('bit_exact_vs_gmp'))
"
tail -2 gpurun_out/bench_${wl}_ct$t.err
done; done
