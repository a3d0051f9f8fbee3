"""One large product spread over the GPUs of a box (BASELINE configs[4] and the sizes below it).

  torchrun --nproc-per-node N scripts/big_sharded.py --log2 24 --depth 16 --w 1 [--steps 2]

Operands are generated on the device (splitmix64 counter generator), every rank holds
them in full and reads only the pieces of its own columns.  The result stays sharded; it is
checked without a CPU product through residues: a*b == r modulo 2^31-1 and modulo 2^61-1, each rank
reducing its own window of result limbs (2^64 is a small power of two modulo both primes, so a
residue is a short weighted sum of per-class limb sums).  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

M31, M61 = (1 << 31) - 1, (1 << 61) - 1


def splitmix64_dev(seed, n, dev, chunk=1 << 26):
    """limb[k] = splitmix64(seed + k) as int64 bit patterns, generated on the device in chunks"""
    out = torch.empty(n, dtype=torch.int64, device=dev)

    def lsr(z, s):
        return (z >> s) & ((1 << (64 - s)) - 1)

    def c64(v):                      # python int -> the int64 with the same bit pattern
        return v - (1 << 64) if v >= (1 << 63) else v

    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        z = torch.arange(lo, hi, dtype=torch.int64, device=dev) + c64(seed & ((1 << 64) - 1))
        z = z * c64(0x9E3779B97F4A7C15) + c64(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * c64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * c64(0x94D049BB133111EB)
        out[lo:hi] = z ^ lsr(z, 31)
    return out


def residue_classes(limbs, first_index, chunk=1 << 26):
    """per-class sums that determine (sum_k limb_k 2^(64 (first_index+k))) mod 2^31-1 and 2^61-1.
    2^64 == 4 (mod 2^31-1): the weight of limb k is 2^(2k mod 31); 2^64 == 8 (mod 2^61-1): 2^(3k mod 61).
    Returns int64 tensors [31] and [61, 2] (the second in two 31-bit halves to stay inside int64)."""
    dev = limbs.device
    s31 = torch.zeros(31, dtype=torch.int64, device=dev)
    s61 = torch.zeros(61, 2, dtype=torch.int64, device=dev)
    m31, m61, lo31 = M31, M61, (1 << 31) - 1
    for lo in range(0, limbs.numel(), chunk):
        x = limbs[lo:lo + chunk]
        k = torch.arange(lo, lo + x.numel(), dtype=torch.int64, device=dev) + first_index
        # x mod 2^31-1 from the three 31/31/2-bit pieces of the unsigned limb
        a0, a1, a2 = x & m31, (x >> 31) & m31, (x >> 62) & 3
        v31 = a0 + a1 + a2              # weights 1, 2^31 == 1, 2^62 == 1
        s31.index_add_(0, k % 31, v31)
        # x mod 2^61-1: low 61 bits + top 3 bits
        v61 = (x & m61) + ((x >> 61) & 7)
        cls = k % 61
        s61[:, 0].index_add_(0, cls, v61 & lo31)
        s61[:, 1].index_add_(0, cls, v61 >> 31)
    return s31, s61


def residues_from_classes(s31, s61):
    s31, s61 = s31.cpu().tolist(), s61.cpu().tolist()
    r31 = sum(v * pow(2, (2 * c) % 31, M31) for c, v in enumerate(s31)) % M31
    r61 = sum((lo + (hi << 31)) * pow(2, (3 * c) % 61, M61) for c, (lo, hi) in enumerate(s61)) % M61
    return r31, r61


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=24, help="operands of 2^LOG2 limbs each")
    ap.add_argument("--depth", type=int, default=16)
    ap.add_argument("--w", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import mpir_fft_b200 as M
    from mpir_fft_b200.sharded import ShardedMul
    M.init(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = 1 << args.log2
    t0 = time.perf_counter()
    a = splitmix64_dev(0x5EED0001, n, dev)
    b = splitmix64_dev(0x5EED0002, n, dev)
    sm = ShardedMul(n, n, args.depth, args.w, cuda=True)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    lay = sm.lay
    for _ in range(args.warmup):
        sm.multiply(a.data_ptr(), b.data_ptr())
    torch.cuda.synchronize()
    times = []
    for _ in range(args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sm.multiply(a.data_ptr(), b.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    # check: residues of the operands (rank 0) against the residues of the sharded result
    ra31, ra61 = residues_from_classes(*residue_classes(a, 0))
    rb31, rb61 = residues_from_classes(*residue_classes(b, 0))
    s31, s61 = residue_classes(sm.out, int(lay.limb_lo))
    if world > 1:
        dist.all_reduce(s31)
        dist.all_reduce(s61)
    rr31, rr61 = residues_from_classes(s31, s61)
    ok = (ra31 * rb31 % M31 == rr31) and (ra61 * rb61 % M61 == rr61)
    if rank == 0:
        ms = min(times)
        print(json.dumps({
            "metric": "new_mpn_mul Mlimb/s", "value": 2 * n / (ms * 1e-3) / 1e6, "unit": "Mlimb/s", "n_gpus": world,
            "ms_per_step": ms, "all_steps_ms": times, "scaling": "strong", "dtype": "u64", "data": "synthetic",
            "config": {"workload": "new_mpn_mul 2^%d x 2^%d limbs, depth %d, w %d, ONE product sharded over %d rank(s)" % (
                           args.log2, args.log2, args.depth, args.w, world),
                       "coefficient_limbs": int(lay.block_limbs) - 2, "all_to_all_bytes_per_rank": int(lay.trunc_rows * lay.ncl * lay.block_limbs * 8)},
            "check": "a*b == r mod 2^31-1 and mod 2^61-1 (result limbs reduced where they live)", "residues_match": bool(ok),
            "setup_s": t_setup, "device_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
    sm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
