"""CPU: the sharded multiplication (csrc/host/smul.c + mpir_fft_b200/sharded.py) end to end on the
CPU-emulated twin of the library, world sizes 1, 2 and 4 over gloo, against GMP.  The emulated
library runs the same host C and the same kernel source; only the collectives differ from the
GPU box (gloo on host memory instead of NCCL on HBM)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")


def _load_emu():
    from mpir_fft_b200._lib import bind
    return bind(C.CDLL(os.path.join(EMU_DIR, "libmpirfft_emu.so"), mode=C.RTLD_LOCAL))


def _run_case(L, case):
    from common import operand, ptr
    from oracle import loader as ORA
    from mpir_fft_b200.sharded import ShardedMul
    n1, n2, depth, w, kind = case
    a, b = operand(kind, n1, 1), operand(kind, n2, 2)
    sm = ShardedMul(n1, n2, depth, w, L=L, cuda=False)
    sm.multiply(ptr(a), ptr(b))
    r = sm.gather_result()
    sm.close()
    return bool(np.array_equal(r, ORA.gmp_mul(a, b)))


def _worker(rank, world, port, cases, q):
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = _load_emu()
    try:
        res = [_run_case(L, c) for c in cases]
    except Exception as e:                       # report instead of hanging the peer
        res = ["%s: %s" % (type(e).__name__, e)]
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


CASES = [
    (3000, 2000, 6, 128, "uniform"),     # l = 128, fused tiles
    (6000, 6000, 6, 256, "ones"),        # l = 256 (cfg2 ring), worst-case carries across the rank windows
    (700, 900, 7, 12, "runs"),           # l = 24: stagewise path
    (8000, 8000, 8, 16, "uniform"),      # bit-shifted twiddles
    (5000, 37, 7, 64, "uniform"),        # lopsided: some ranks own no result limbs
    (30000, 28000, 6, 1024, "ones"),     # l = 1024: stage-per-launch transforms, pointwise products through the
                                         # batched negacyclic mulmod (the route of the 2^24 .. 2^30-limb products)
]


@pytest.fixture(scope="module")
def emu_built():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])


def test_sharded_world1(emu_built):
    L = _load_emu()
    for c in CASES:
        assert _run_case(L, c), c


@pytest.mark.parametrize("world,port", [(2, 29541), (4, 29543)])
def test_sharded_gloo(emu_built, world, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, CASES, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank, r in res:
        assert r == [True] * len(CASES), (rank, r)
