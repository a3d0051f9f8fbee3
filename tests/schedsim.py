"""Test helper: a big-integer interpreter for the op lists that the C host side generates
(mpir_fft_b200/csrc/host/sched.c).  It executes exactly the descriptors the GPU executes, with
Python ints mod p = 2^NW + 1, so the schedule logic (recursion walk, slot ping-pong, stage
assignment, twist exponents) is verified on the CPU against the compiled reference.
"""
import ctypes as C
import numpy as np
from mpir_fft_b200 import lib

NONE = 0xFFFFFFFF
KINDS = dict(FFT=0, FFT_TRUNC=1, FFT_TRUNC1=2, IFFT=3, IFFT_TRUNC=4, IFFT_TRUNC1=5,
             FFT_NEGACYCLIC=6, IFFT_NEGACYCLIC=7)


class Op(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("inA", "inB", "outS", "outT", "eSA", "eSB", "eTA", "eTB",
                                           "cSA", "cSB", "cTA", "cTB")] + \
               [(k, C.c_int8) for k in ("sSA", "sSB", "sTA", "sTB")] + [("stage", C.c_uint32)] + \
               [(k, C.c_uint32) for k in ("pA", "pB", "pS", "pT", "pstage", "kind", "kparam")]


class Sched(C.Structure):
    _fields_ = [("S", C.c_uint32), ("NW", C.c_uint64), ("M2", C.c_uint64),
                ("slot", C.POINTER(C.c_uint32)), ("wr_stage", C.POINTER(C.c_uint32)),
                ("rd_stage", C.POINTER(C.c_uint32)), ("ops", C.POINTER(Op)),
                ("nops", C.c_size_t), ("cap", C.c_size_t), ("nstages", C.c_uint32),
                ("stage_off", C.POINTER(C.c_uint32)), ("phys", C.POINTER(C.c_uint32)),
                ("pwr_stage", C.POINTER(C.c_uint32)), ("prd_stage", C.POINTER(C.c_uint32)),
                ("npstages", C.c_uint32)]


def _setup():
    L = lib()
    L.mfft_sched_new.restype = C.POINTER(Sched)
    L.mfft_sched_new.argtypes = [C.c_uint32, C.c_uint64]
    L.mfft_sched_free.argtypes = [C.POINTER(Sched)]
    L.mfft_sched_emit.restype = C.c_int
    L.mfft_sched_emit.argtypes = [C.POINTER(Sched), C.c_int, C.c_uint32, C.c_uint32] + [C.c_uint64] * 6
    L.mfft_sched_revbin.argtypes = [C.POINTER(Sched), C.c_uint32, C.c_uint32, C.c_uint32]
    L.mfft_sched_finish.restype = C.c_int
    L.mfft_sched_finish.argtypes = [C.POINTER(Sched)]
    return L


def block_to_int(blk, l):
    """(l+1)-limb two's-complement block -> residue mod 2^(64 l)+1 (mul_fft.c:3677-3697)."""
    body = int.from_bytes(np.asarray(blk[:l], dtype=np.uint64).tobytes(), "little")
    top = int(np.asarray(blk[l:l + 1], dtype=np.uint64).view(np.int64)[0])
    return (body + (top << (64 * l))) % ((1 << (64 * l)) + 1)


def int_to_block(v, l):
    """canonical residue -> normalised block (top limb 0, or 1 with zero body)"""
    out = np.zeros(l + 1, dtype=np.uint64)
    if v == 1 << (64 * l):
        out[l] = 1
    else:
        out[:l] = np.frombuffer(v.to_bytes(8 * l, "little"), dtype=np.uint64)
    return out


class SimSched:
    def __init__(self, S, NW):
        self.L = _setup()
        self.S, self.NW = S, NW
        self.p = (1 << NW) + 1
        self.h = self.L.mfft_sched_new(S, NW)

    def emit(self, kind, p0, is_, n, w, ws=0, r=0, rs=0, trunc=0):
        rc = self.L.mfft_sched_emit(self.h, KINDS[kind], p0, is_, n, w, ws, r, rs, trunc)
        if rc != 0:
            raise ValueError("mfft_sched_emit rejected the parameters")

    def revbin(self, p0, is_, bits):
        self.L.mfft_sched_revbin(self.h, p0, is_, bits)

    def finish(self):
        assert self.L.mfft_sched_finish(self.h) == 0

    def ops(self):
        s = self.h.contents
        return [s.ops[i] for i in range(s.nops)]

    def slots(self):
        s = self.h.contents
        return [s.slot[i] for i in range(self.S)]

    def nstages(self):
        return self.h.contents.nstages

    def run(self, values, col=0, check_stages=True):
        """values: list of S residues (logical positions, all in half 0). Returns the residue of
        every logical position after executing the ops stage by stage."""
        s = self.h.contents
        mem = {i: v for i, v in enumerate(values)}
        M2, p = 2 * self.NW, self.p
        for st in range(1, s.nstages + 1):
            lo, hi = s.stage_off[st - 1], s.stage_off[st]
            written, reads, pending = set(), set(), {}
            for i in range(lo, hi):
                o = s.ops[i]
                assert o.stage == st
                A = mem[o.inA]
                B = mem[o.inB] if o.inB != NONE else 0
                reads.add(o.inA)
                if o.inB != NONE:
                    reads.add(o.inB)
                vS = (o.sSA * A * pow(2, (o.eSA + col * o.cSA) % M2, p) +
                      o.sSB * B * pow(2, (o.eSB + col * o.cSB) % M2, p)) % p
                assert o.outS not in written and o.outS not in pending
                pending[o.outS] = vS
                if o.outT != NONE:
                    vT = (o.sTA * A * pow(2, (o.eTA + col * o.cTA) % M2, p) +
                          o.sTB * B * pow(2, (o.eTB + col * o.cTB) % M2, p)) % p
                    assert o.outT not in pending
                    pending[o.outT] = vT
            if check_stages:     # ops of one stage must be independent: nobody reads what another writes
                assert not (reads & set(pending)), "stage %d has a read/write overlap" % st
            mem.update(pending)
        return [mem[sl] for sl in self.slots()]

    def close(self):
        if self.h:
            self.L.mfft_sched_free(self.h)
            self.h = None


# ---------------------------------------------------------------------------------------------
# whole-MFA simulation: the same addressing arithmetic as k_run_stage / k_finalize
# ---------------------------------------------------------------------------------------------
class Batch(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("base", "parity", "col", "pad")]


class Move(C.Structure):
    _fields_ = [("src_slot", C.c_uint32), ("dst_pos", C.c_uint32)]


def _block_index(S, slot_stride, half_blocks, slot, b):
    return ((slot // S) ^ b.parity) * half_blocks + b.base + (slot % S) * slot_stride


def _run_pass(mem, sched, slot_stride, half_blocks, batches, p, M2):
    s = sched.contents
    S = s.S
    for st in range(1, s.nstages + 1):
        pending = {}
        for i in range(s.stage_off[st - 1], s.stage_off[st]):
            o = s.ops[i]
            for b in batches:
                A = mem.get(_block_index(S, slot_stride, half_blocks, o.inA, b), 0)
                B = mem.get(_block_index(S, slot_stride, half_blocks, o.inB, b), 0) if o.inB != NONE else 0
                vS = (o.sSA * A * pow(2, (o.eSA + b.col * o.cSA) % M2, p) +
                      o.sSB * B * pow(2, (o.eSB + b.col * o.cSB) % M2, p)) % p
                k = _block_index(S, slot_stride, half_blocks, o.outS, b)
                assert k not in pending
                pending[k] = vS
                if o.outT != NONE:
                    vT = (o.sTA * A * pow(2, (o.eTA + b.col * o.cTA) % M2, p) +
                          o.sTB * B * pow(2, (o.eTB + b.col * o.cTB) % M2, p)) % p
                    k = _block_index(S, slot_stride, half_blocks, o.outT, b)
                    assert k not in pending
                    pending[k] = vT
        mem.update(pending)


def simulate_mfa(values, inverse, n, w, n1, trunc, shift=0):
    """values: residues of the 2n input blocks in reference order.  Returns {dst block: residue}
    for every block the device transform would write (mfft_mfa_exec)."""
    L = _setup()
    L.mfft_mfa_debug_new.restype = C.c_void_p
    L.mfft_mfa_debug_new.argtypes = [C.c_int] + [C.c_uint64] * 4
    L.mfft_mfa_debug_free.argtypes = [C.c_void_p]
    L.mfft_mfa_debug_sched.restype = C.POINTER(Sched)
    L.mfft_mfa_debug_sched.argtypes = [C.c_void_p, C.c_int]
    L.mfft_mfa_debug_batch.restype = C.POINTER(Batch)
    L.mfft_mfa_debug_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32)]
    L.mfft_mfa_debug_moves.restype = C.POINTER(Move)
    L.mfft_mfa_debug_moves.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.mfft_mfa_debug_dst_base.restype = C.POINTER(C.c_uint32)
    L.mfft_mfa_debug_dst_base.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    m = L.mfft_mfa_debug_new(int(inverse), n, w, n1, trunc)
    if not m:
        raise ValueError("illegal MFA parameters")
    try:
        NW = n * w
        p, M2, N, n2 = (1 << NW) + 1, 2 * NW, 2 * n, 2 * n // n1
        cnt = C.c_uint32()
        cs, rs = L.mfft_mfa_debug_sched(m, 0), L.mfft_mfa_debug_sched(m, 1)
        cb = L.mfft_mfa_debug_batch(m, 0, C.byref(cnt)); colb = [cb[i] for i in range(cnt.value)]
        rb = L.mfft_mfa_debug_batch(m, 1, C.byref(cnt)); rowb = [rb[i] for i in range(cnt.value)]
        stride = C.c_uint32()
        mv = L.mfft_mfa_debug_moves(m, C.byref(cnt), C.byref(stride)); moves = [mv[i] for i in range(cnt.value)]
        db = L.mfft_mfa_debug_dst_base(m, C.byref(cnt)); dstb = [db[i] for i in range(cnt.value)]
        mem = {i: v for i, v in enumerate(values)}
        if not inverse:
            _run_pass(mem, cs, n1, N, colb, p, M2)
            _run_pass(mem, rs, 1, N, rowb, p, M2)
            S, sstr, fb = n1, 1, rowb
        else:
            _run_pass(mem, rs, 1, N, rowb, p, M2)
            _run_pass(mem, cs, n1, N, colb, p, M2)
            S, sstr, fb = n2, n1, colb
        out = {}
        for mvi in moves:
            for bi, b in enumerate(fb):
                src = _block_index(S, sstr, N, mvi.src_slot, b)
                out[dstb[bi] + mvi.dst_pos * stride.value] = mem[src] * pow(2, shift % M2, p) % p
        return out
    finally:
        L.mfft_mfa_debug_free(m)


def simulate_mfa_sqrt2(values, inverse, n, w, n1, trunc, shift=0):
    """the same for the sqrt2 MFA (4n blocks; two column classes when w is odd): {dst block: residue}"""
    L = _setup()
    L.mfft_mfa_debug_new_sqrt2.restype = C.c_void_p
    L.mfft_mfa_debug_new_sqrt2.argtypes = [C.c_int] + [C.c_uint64] * 4
    L.mfft_mfa_debug_free.argtypes = [C.c_void_p]
    L.mfft_mfa_debug_sched.restype = C.POINTER(Sched)
    L.mfft_mfa_debug_sched.argtypes = [C.c_void_p, C.c_int]
    L.mfft_mfa_debug_batch.restype = C.POINTER(Batch)
    L.mfft_mfa_debug_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32)]
    L.mfft_mfa_debug_moves.restype = C.POINTER(Move)
    L.mfft_mfa_debug_moves.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.mfft_mfa_debug_dst_base.restype = C.POINTER(C.c_uint32)
    L.mfft_mfa_debug_dst_base.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.mfft_mfa_debug_dst_base2.restype = C.POINTER(C.c_uint32)
    L.mfft_mfa_debug_dst_base2.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    m = L.mfft_mfa_debug_new_sqrt2(int(inverse), n, w, n1, trunc)
    if not m:
        raise ValueError("illegal sqrt2 MFA parameters")
    try:
        NW = n * w
        p, M2, N, n2 = (1 << NW) + 1, 2 * NW, 4 * n, 2 * n // n1
        cnt = C.c_uint32()
        classes = [0, 2] if (w & 1) else [0]
        cs = [L.mfft_mfa_debug_sched(m, k) for k in classes]
        colb = []
        for k in classes:
            cb = L.mfft_mfa_debug_batch(m, k, C.byref(cnt))
            colb.append([cb[i] for i in range(cnt.value)])
        rs = L.mfft_mfa_debug_sched(m, 1)
        rb = L.mfft_mfa_debug_batch(m, 1, C.byref(cnt)); rowb = [rb[i] for i in range(cnt.value)]
        stride = C.c_uint32()
        mv = L.mfft_mfa_debug_moves(m, C.byref(cnt), C.byref(stride)); moves = [mv[i] for i in range(cnt.value)]
        db = L.mfft_mfa_debug_dst_base(m, C.byref(cnt)); dstb = [[db[i] for i in range(cnt.value)]]
        if inverse and len(classes) == 2:
            db2 = L.mfft_mfa_debug_dst_base2(m, C.byref(cnt)); dstb.append([db2[i] for i in range(cnt.value)])
        mem = {i: v for i, v in enumerate(values)}
        out = {}
        if not inverse:
            for c, b in zip(cs, colb):
                _run_pass(mem, c, n1, N, b, p, M2)
            _run_pass(mem, rs, 1, N, rowb, p, M2)
            for mvi in moves:
                for bi, b in enumerate(rowb):
                    src = _block_index(n1, 1, N, mvi.src_slot, b)
                    out[dstb[0][bi] + mvi.dst_pos * stride.value] = mem[src] * pow(2, shift % M2, p) % p
        else:
            _run_pass(mem, rs, 1, N, rowb, p, M2)
            for c, b, d in zip(cs, colb, dstb):
                _run_pass(mem, c, n1, N, b, p, M2)
                for mvi in moves:
                    for bi, be in enumerate(b):
                        src = _block_index(2 * n2, n1, N, mvi.src_slot, be)
                        out[d[bi] + mvi.dst_pos * stride.value] = mem[src] * pow(2, shift % M2, p) % p
        return out
    finally:
        L.mfft_mfa_debug_free(m)
