#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
MPIRFFT_BIG_INPLACE=$v MPIRFFT_VERBOSE=$v timeout 600 python - > gpurun_out/big_$v.log 2> gpurun_out/big_$v.err <<PY
import json, torch, sys
sys.path.insert(0, ".")
import bench, mpir_fft_b200 as M
torch.cuda.set_device(0); M.init(0)
import torch.distributed as dist
for lg in (26, 24):
    rec = bench.sharded_leg(torch, dist, M, 0, 1, lg, steps=2, peak_gbs=6557.1)
    print(json.dumps({k: rec[k] for k in ("workload", "ms_per_step", "phases_ms", "bit_exact")}))
PY
echo "BIG_INPLACE=$v rc=$?"; cat gpurun_out/big_$v.log | cut -c1-700
done
MPIRFFT_BIG_INPLACE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/big_launches.csv python scripts/big_one.py 26 1 > gpurun_out/big_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/big_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hd=rows[h]; iN=hd.index("Kernel Name"); iV=hd.index("Metric Value"); iG=hd.index("Grid Size") if "Grid Size" in hd else None
for r in rows[h+2:]:
    if len(r)>iV: print(r[iN][:40].ljust(40), r[iG] if iG is not None else "", r[iV])
PY
