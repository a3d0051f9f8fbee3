"""developer aid: one sharded leg of bench.py under torchrun:  torchrun ... scripts/leg.py LOG2"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import bench, mpir_fft_b200 as M
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); M.init(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rec = bench.sharded_leg(torch, dist, M, rank, world, int(sys.argv[1]), steps=2, baseline_1gpu=False, peak_gbs=6557.1)
if rank == 0:
    print(json.dumps({k: rec[k] for k in ("workload", "ms_per_step", "phases_ms", "bit_exact")}))
if world > 1:
    dist.destroy_process_group()
