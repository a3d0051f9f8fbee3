"""new_mpn_mul with the MFA sharded over the GPUs of one box: one process per GPU.

The library (csrc/host/smul.c) runs the LOCAL phases on plan-owned HBM buffers; this module only
issues the exchange steps between them with torch.distributed -- NCCL over NVLink on the GPUs,
gloo in the CPU tests (where the library handle passed in is the CPU-emulated twin and the
"device" pointers are host memory).  Phase order per product (smul.c header):

    0 fwd_cols(i1) A 1 fwd_rows | 0 fwd_cols(i2) A 1 fwd_rows | 2 pointwise
    3 inv_rows B 4 inv_cols C 5 unpack H 6 recombine, then the carry hand-off rank 0 -> world-1

A, C: all_to_all split by live rows; B: all_to_all split by columns; H: all_gather of every
rank's last `halo_send` coefficients (a rank's first result limbs are reached by them).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import lib as _default_lib
from .dist_util import shard_units


class SmulLayout(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in
                ("block_limbs", "ncl", "nrl", "r0", "trunc_rows", "n1cols", "halo", "halo_send")] + \
               [(k, C.c_uint64) for k in ("limb_lo", "limb_hi", "send_limbs", "recv_limbs", "work_limbs")] + \
               [(k, C.c_void_p) for k in ("send", "recv", "work", "unp", "out")]


def bind_smul(L):
    vp, i64, u64, i32, u32 = C.c_void_p, C.c_long, C.c_uint64, C.c_int, C.c_uint
    L.mpirfft_smul_plan_create.restype, L.mpirfft_smul_plan_create.argtypes = i32, [C.POINTER(vp), i64, i64, u64, u64, i32, i32]
    L.mpirfft_smul_plan_destroy.restype, L.mpirfft_smul_plan_destroy.argtypes = None, [vp]
    L.mpirfft_smul_info.restype, L.mpirfft_smul_info.argtypes = i32, [vp, C.POINTER(SmulLayout)]
    L.mpirfft_smul_phase.restype, L.mpirfft_smul_phase.argtypes = i32, [vp, i32, i32, vp, vp]
    L.mpirfft_smul_carry.restype, L.mpirfft_smul_carry.argtypes = i32, [vp, u32, C.POINTER(u32), vp]
    L.mpirfft_smul_tail_blocks.restype, L.mpirfft_smul_tail_blocks.argtypes = vp, [vp, u32]
    return L


class _DevView:
    """a raw device pointer dressed up for torch.as_tensor (CUDA array interface, int64 limbs)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def _view(ptr, n, cuda):
    """int64 tensor of n limbs over library-owned memory (no copy)"""
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device="cuda" if cuda else "cpu")
    if cuda:
        return torch.as_tensor(_DevView(ptr, n), device="cuda")
    arr = np.ctypeslib.as_array(C.cast(C.c_void_p(ptr), C.POINTER(C.c_int64)), shape=(n,))
    return torch.from_numpy(arr)


class ShardedMul:
    """One rank's part of a sharded product n1 x n2 limbs at (depth, w).

    `L`: the bound library (default: the product library); `group`: the process group of the box;
    `single`: ignore the process group and run the whole product on this rank.
    Operands are given to every rank in full (device pointers, or numpy arrays for the emulated
    library); a rank reads only the pieces of its own columns."""

    def __init__(self, n1, n2, depth, w, L=None, group=None, cuda=None, single=False):
        self.L = bind_smul(L if L is not None else _default_lib())
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if single:               # the whole product on this rank, no collectives (1-GPU baseline inside a multi-rank job)
            self.rank, self.world = 0, 1
        self.cuda = torch.cuda.is_available() if cuda is None else cuda
        self.n1, self.n2 = n1, n2
        h = C.c_void_p()
        rc = self.L.mpirfft_smul_plan_create(C.byref(h), n1, n2, depth, w, self.rank, self.world)
        if rc != 0:
            raise RuntimeError("mpirfft_smul_plan_create failed (%d): %s" % (rc, self.L.mpirfft_last_error().decode()))
        self.h = h
        lay = SmulLayout()
        self.L.mpirfft_smul_info(self.h, C.byref(lay))
        self.lay = lay
        P = lay.block_limbs
        self.P = P
        self.send = _view(lay.send, lay.send_limbs, self.cuda)
        self.recv = _view(lay.recv, lay.recv_limbs, self.cuda)
        self.work = _view(lay.work, lay.work_limbs, self.cuda)
        self.out = _view(lay.out, lay.limb_hi - lay.limb_lo, self.cuda)
        rows = [shard_units(lay.trunc_rows, r, self.world) for r in range(self.world)]
        self.rows_of = [hi - lo for lo, hi in rows]
        # A / C: to rank h go its rows of my columns; from rank g come my rows of its columns
        self.a_in = [n * lay.ncl * P for n in self.rows_of]
        self.a_out = [lay.nrl * lay.ncl * P] * self.world
        # B: to rank g go its columns of my rows; from rank h come its rows of my columns
        self.b_in = [lay.nrl * lay.ncl * P] * self.world
        self.b_out = [n * lay.ncl * P for n in self.rows_of]
        self.tr_limbs = lay.trunc_rows * lay.ncl * P
        self.halo_buf = None

    def close(self):
        if self.h:
            self.L.mpirfft_smul_plan_destroy(self.h)
            self.h = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream) if self.cuda else None

    def _phase(self, ph, which=0, d_in=None):
        rc = self.L.mpirfft_smul_phase(self.h, ph, which, d_in, self._stream())
        if rc != 0:
            raise RuntimeError("mpirfft_smul_phase %d failed (%d): %s" % (ph, rc, self.L.mpirfft_last_error().decode()))

    def _a2a(self, out, inp, out_split, in_split):
        if self.world == 1:
            out[:inp.numel()].copy_(inp)
            return
        dist.all_to_all_single(out, inp, out_split, in_split, group=self.group)

    def _mark(self, timing, name):
        """phase boundary for the per-phase CUDA-event timing (bench.py); nothing when timing is off"""
        if timing is None or not self.cuda:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        timing.append((name, ev))

    def multiply(self, d_i1, d_i2, timing=None):
        """run the product; afterwards self.out holds limbs [limb_lo, limb_hi) of it.
        timing: a list that receives (phase name, CUDA event recorded when the phase has been issued)
        pairs on the current stream (the NCCL collectives are ordered against it)"""
        lay, P = self.lay, self.P
        self._mark(timing, "start")
        for which, d_in in ((0, d_i1), (1, d_i2)):
            self._phase(0, which, d_in)
            self._mark(timing, "fwd_cols")
            self._a2a(self.recv[:sum(self.a_out)], self.send[:self.tr_limbs], self.a_out, self.a_in)
            self._mark(timing, "all_to_all")
            self._phase(1, which)
            self._mark(timing, "fwd_rows")
        self._phase(2)
        self._mark(timing, "pointwise")
        self._phase(3)
        self._mark(timing, "inv_rows")
        self._a2a(self.work[:self.tr_limbs], self.send[:sum(self.b_in)], self.b_out, self.b_in)
        self._mark(timing, "all_to_all")
        self._phase(4)
        self._mark(timing, "inv_cols")
        self._a2a(self.recv[:sum(self.a_out)], self.send[:self.tr_limbs], self.a_out, self.a_in)
        self._mark(timing, "all_to_all")
        self._phase(5)
        if self.world > 1:
            # halo: my last `halo_send` coefficients go to the next rank only, straight from the
            # plan's row buffer into the successor's halo slots (no staging copy)
            hs = lay.halo_send
            n = hs * P
            tail = _view(self.L.mpirfft_smul_tail_blocks(self.h, hs), n, self.cuda)
            unp = _view(lay.unp, n, self.cuda)
            last = self.rank + 1 == self.world
            in_split = [n if (g == self.rank + 1) else 0 for g in range(self.world)]
            out_split = [n if (g == self.rank - 1) else 0 for g in range(self.world)]
            dist.all_to_all_single(unp[:0 if self.rank == 0 else n], tail[:0 if last else n], out_split, in_split, group=self.group)
        self._phase(6)
        self._mark(timing, "recombine")
        self._carry_handoff()
        self._mark(timing, "carry")
        return self.out

    def _carry(self, carry_in):
        co = C.c_uint(0)
        rc = self.L.mpirfft_smul_carry(self.h, int(carry_in), C.byref(co), self._stream())
        if rc != 0:
            raise RuntimeError("mpirfft_smul_carry failed (%d)" % rc)
        return int(co.value)

    def _carry_handoff(self):
        """carries between the rank windows (FFT_combine_bits' carry to the end of r1, mul_fft.c:207-267).
        Every rank publishes the carry that left its window (one all-gather); rank r adds the carry of
        rank r-1.  Such an addition ripples out of a whole window only if every limb of it is all ones,
        so a second all-gather of the ripples almost always finds zeros; otherwise the ripples are handed
        on the same way until none is left (additions commute, so the result is exact)."""
        own = self._carry(0)                      # what left my window in phase 6 (synchronises my stream)
        if self.world == 1:
            return
        dev = self.out.device
        pend = own
        for _ in range(self.world + 1):
            mine = torch.tensor([pend], dtype=torch.int64, device=dev)
            every = torch.empty(self.world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(every, mine, group=self.group)
            every = every.cpu().tolist()
            if not any(every[:-1]):               # what leaves the last window is beyond n1 + n2 limbs: always 0
                return
            cin = every[self.rank - 1] if self.rank > 0 else 0
            pend = (self._carry(cin) - own) if cin else 0     # the ripple of this addition only
        raise RuntimeError("carry hand-off did not settle")

    def gather_result(self):
        """the whole product on every rank (testing / small sizes): limbs as a uint64 numpy array"""
        mine = self.out.cpu().numpy().view(np.uint64)
        if self.world == 1:
            return mine.copy()
        parts = [None] * self.world
        dist.all_gather_object(parts, mine, group=self.group)
        return np.concatenate(parts)
