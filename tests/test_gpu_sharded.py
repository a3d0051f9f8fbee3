"""GPU: the sharded multiplication over NCCL, one process per GPU, bit-exact vs GMP.
World size = min(visible GPUs, 8) restricted to a power of two; with one GPU the same code runs
as world 1 (all exchanges degenerate to local copies)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    (1 << 20, 1 << 20, 14, 1, "uniform"),       # cfg2
    (3000000, 1700000, 14, 2, "uniform"),       # cfg3, trunc > n branch, l = 512
    (1 << 18, 1 << 18, 13, 1, "ones"),          # odd depth, worst-case carries across rank windows
    (1 << 22, 1000, 15, 1, "runs"),             # lopsided
    (1 << 22, 1 << 22, 16, 1, "ones"),          # ring of 1024 limbs: carry-save stage kernel + transform-based pointwise
    (6000000, 2000000, 16, 1, "runs"),          # the same route, truncated and lopsided
]


def _worker(rank, world, port, cases, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import mpir_fft_b200 as M
    from mpir_fft_b200.sharded import ShardedMul
    from common import operand
    from oracle import loader as ORA
    res = []
    try:
        torch.cuda.set_device(rank)
        M.init(rank)
        if world > 1:
            os.environ["MASTER_ADDR"] = "127.0.0.1"
            os.environ["MASTER_PORT"] = str(port)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        for n1, n2, depth, w, kind in cases:
            a, b = operand(kind, n1, 1), operand(kind, n2, 2)
            da = torch.from_numpy(a.view(np.int64)).cuda()
            db = torch.from_numpy(b.view(np.int64)).cuda()
            sm = ShardedMul(n1, n2, depth, w, cuda=True)
            sm.multiply(da.data_ptr(), db.data_ptr())
            torch.cuda.synchronize()
            r = sm.gather_result()
            sm.close()
            res.append(bool(np.array_equal(r, ORA.gmp_mul(a, b))) if rank == 0 else True)
    except Exception as e:                      # report instead of hanging the peers
        import traceback
        res.append("%s: %s\n%s" % (type(e).__name__, e, traceback.format_exc()))
    q.put((rank, res))
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass


@pytest.mark.gpu
def test_sharded_nccl():
    ngpu = torch.cuda.device_count()
    world = 1
    while world * 2 <= min(ngpu, 8):
        world *= 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29561, CASES, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    for rank, r in res:
        assert r == [True] * len(CASES), (rank, world, r)
