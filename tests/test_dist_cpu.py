"""CPU, world_size 2 over gloo: the rank partition and the max-over-ranks timing reduction the
multi-GPU bench uses."""
import os

import torch.distributed as dist
import torch.multiprocessing as mp

from mpir_fft_b200.dist_util import shard_units, max_over_ranks, sum_over_ranks


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_units(130, rank, world)            # the 130 live rows of cfg2
    mx = max_over_ranks(10.0 + rank)
    sm = sum_over_ranks(hi - lo)
    q.put((rank, lo, hi, mx, sm))
    dist.destroy_process_group()


def test_shard_and_reduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29533, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [(r[1], r[2]) for r in res] == [(0, 65), (65, 130)]
    assert all(r[3] == 11.0 for r in res) and all(r[4] == 130.0 for r in res)


def test_shard_units_cover_everything():
    for total in (0, 1, 7, 130, 1026):
        for world in (1, 2, 3, 8):
            parts = [shard_units(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
