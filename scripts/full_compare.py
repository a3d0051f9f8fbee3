"""Limb-for-limb comparison of ONE large sharded product with GMP mpn_mul on the host.

  python scripts/full_compare.py --log2 26                      (one GPU, the whole product on it)
  torchrun --nproc-per-node N scripts/full_compare.py --log2 26 (the product sharded over N ranks)

The operands are the bench's synthetic limbs (splitmix64 counter generator).  Every rank copies its
window of the result to the host; rank 0 assembles the product, multiplies the operands once more with
GMP (test infrastructure: oracle/loader.py, libgmp of the image) and compares every limb.  Also
prints the residue fingerprint of mpir_fft_b200/residues.py so that cheap checks of later runs can be
tied to this full comparison.  One JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=26)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--w", type=int, default=1)
    args = ap.parse_args()
    import bench
    depth, w = (args.depth, args.w) if args.depth else bench.SHARDED_PARAMS[args.log2]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import mpir_fft_b200 as M
    from mpir_fft_b200.sharded import ShardedMul
    from mpir_fft_b200 import residues as RES
    M.init(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = 1 << args.log2
    a = bench.splitmix64_dev(torch, 0x5EED0001, n, dev)
    b = bench.splitmix64_dev(torch, 0x5EED0002, n, dev)
    sm = ShardedMul(n, n, depth, w, cuda=True)
    sm.multiply(a.data_ptr(), b.data_ptr())
    torch.cuda.synchronize()
    mine = sm.out.cpu()
    fp = RES.residues(sm.out, int(sm.lay.limb_lo))
    if world > 1:
        t = torch.tensor(fp, dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        fp = [int(v) % p for v, p in zip(t.cpu().tolist(), RES.PRIMES)]
        sizes = [None] * world
        dist.all_gather_object(sizes, int(mine.numel()))
        if rank == 0:
            parts = [mine]
            for g in range(1, world):
                buf = torch.empty(sizes[g], dtype=torch.int64, device=dev)
                dist.recv(buf, src=g)
                parts.append(buf.cpu())
            mine = torch.cat(parts)
        else:
            dist.send(sm.out.contiguous(), dst=0)
    if rank == 0:
        from oracle import loader as oracle
        ha, hb = a.cpu().numpy().view(np.uint64), b.cpu().numpy().view(np.uint64)
        t0 = time.perf_counter()
        want = oracle.gmp_mul(ha, hb)
        t_gmp = time.perf_counter() - t0
        got = mine.numpy().view(np.uint64)
        same = bool(got.shape == want.shape and np.array_equal(got, want))
        first_bad = None if same else int(np.flatnonzero(got[:min(len(got), len(want))] != want[:min(len(got), len(want))])[:1].tolist()[0]) if len(got) else -1
        ra, rb = RES.residues(a, 0), RES.residues(b, 0)
        print(json.dumps({
            "check": "full compare against GMP mpn_mul, every limb",
            "workload": "new_mpn_mul 2^%d x 2^%d limbs, depth %d, w %d, sharded over %d rank(s)" % (args.log2, args.log2, depth, w, world),
            "limbs_compared": int(want.size), "bit_exact": same, "first_differing_limb": first_bad,
            "gmp_mpn_mul_seconds_1_thread": t_gmp,
            "residue_fingerprint": fp, "residues_match": RES.product_matches(ra, rb, fp)}), flush=True)
    sm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
