"""CPU: the *actual kernel source* (csrc/cuda/mfft_kernels.cu) compiled by g++ against the SIMT
emulator in tests/emu and driven through the same C ABI, compared with GMP / big integers.
This is a check of kernel arithmetic and host logic on a machine without a GPU; the emulated
library is test infrastructure and is never loaded by the product package."""
import ctypes as C
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from oracle import loader as L
from common import operand, ptr, block_to_int, int_to_block
from mpir_fft_b200._lib import bind

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    return bind(C.CDLL(os.path.join(EMU_DIR, "libmpirfft_emu.so"), mode=C.RTLD_LOCAL))


@pytest.mark.parametrize("case", [
    (1, 1, 6, 1, "uniform"), (20, 13, 6, 1, "uniform"), (40, 40, 6, 2, "ones"),      # l = 1, 2: generic paths
    (1500, 1500, 6, 64, "ones"), (3000, 2000, 6, 128, "uniform"),                   # l = 64, 128
    (6000, 6000, 6, 256, "ones"), (6000, 10, 6, 256, "runs"),                       # l = 256 (cfg2 ring)
    (12000, 9000, 6, 512, "uniform"), (700, 900, 7, 12, "runs"), (3000, 3000, 7, 96, "ones"),
    (8000, 8000, 8, 16, "uniform"), (3000, 5000, 9, 8, "ones"),                     # fused path, bit-shifted twiddles
])
def test_emulated_new_mpn_mul(emu, case):
    n1, n2, depth, w, kind = case
    a, b = operand(kind, n1, 1), operand(kind, n2, 2)
    r = np.zeros(n1 + n2, dtype=np.uint64)
    emu.new_mpn_mul(ptr(r), ptr(a), n1, ptr(b), n2, depth, w)
    assert np.array_equal(r, L.gmp_mul(a, b))


@pytest.mark.parametrize("case", [(2000, 1500, 6, 1024, "uniform"), (5000, 6000, 5, 2048, "runs")])
def test_emulated_new_mpn_mul_big_ring(emu, case, monkeypatch):
    """Coefficient rings above 512 limbs: new_mpn_mul runs the sharded plan on one rank (mul.c: exec_big --
    sliced in-place transforms, pointwise products through the recursion); MPIRFFT_BIG_PLAN=0 keeps the
    layer-per-launch path with the schoolbook pointwise kernel.  Both must equal GMP."""
    n1, n2, depth, w, kind = case
    a, b = operand(kind, n1, 1), operand(kind, n2, 2)
    want = L.gmp_mul(a, b)
    r = np.zeros(n1 + n2, dtype=np.uint64)
    emu.new_mpn_mul(ptr(r), ptr(a), n1, ptr(b), n2, depth, w)
    assert np.array_equal(r, want)
    if w == 1024:
        monkeypatch.setenv("MPIRFFT_BIG_PLAN", "0")
        a2 = operand(kind, n1 - 1, 1)                      # another plan-cache key: the switch is read at plan creation
        r2 = np.zeros(n1 - 1 + n2, dtype=np.uint64)
        emu.new_mpn_mul(ptr(r2), ptr(a2), n1 - 1, ptr(b), n2, depth, w)
        assert np.array_equal(r2, L.gmp_mul(a2, b))


@pytest.mark.parametrize("case", [
    # (n1, n2, depth, w, operands): 2^(depth+1) < j1+j2-1 <= 2^(depth+2)
    (50, 41, 6, 1, "ones"),                                             # l = 1: stage-per-launch path, odd w (sqrt2 proper)
    (3500, 4000, 6, 64, "ones"),                                        # l = 64, even w: fused tiles
    (5000, 6000, 6, 128, "runs"), (10000, 14000, 6, 256, "ones"),       # l = 128, 256
    (3000, 2500, 8, 3, "uniform"),                                      # l = 12: odd w, two column classes
    (12000, 12000, 8, 16, "uniform"),                                   # l = 64 with bit-shifted twiddles
    (13000, 12000, 8, 17, "runs"),                                      # odd w, l = 68 (stagewise)
])
def test_emulated_new_mpn_mul6(emu, case):
    """new_mpn_mul6 (mul_fft.c:3573): the product through the sqrt2 transforms of length 4n"""
    n1, n2, depth, w, kind = case
    a, b = operand(kind, n1, 1), operand(kind, n2, 2)
    r = np.zeros(n1 + n2, dtype=np.uint64)
    emu.new_mpn_mul6(ptr(r), ptr(a), n1, ptr(b), n2, depth, w)
    assert np.array_equal(r, L.gmp_mul(a, b))


def test_emulated_new_mpn_mul6_vs_golden(emu):
    """new_mpn_mul6 products of the compiled reference, committed as fixtures (tests/golden)"""
    import hashlib
    from golden import make_golden as G
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))
    for n1, n2, depth, w, kind in G.PRODUCTS6[:3]:
        a, b = operand(kind, n1, 0x5EED0001), operand(kind, n2, 0x5EED0002)
        r = np.zeros(n1 + n2, dtype=np.uint64)
        emu.new_mpn_mul6(ptr(r), ptr(a), n1, ptr(b), n2, depth, w)
        key = "P6_%d_%d_%d_%d_%s" % (n1, n2, depth, w, kind)
        if key in gold:
            assert np.array_equal(r, gold[key]), key
        else:
            assert hashlib.sha256(r.tobytes()).digest() == gold[key + "_sha256"].tobytes(), key


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("case", [
    (1500, 1500, 6, 64, "ones"), (3000, 2000, 6, 128, "uniform"), (3000, 3000, 6, 128, "ones"),
    (6000, 6000, 6, 256, "ones"), (6000, 6000, 6, 256, "uniform"), (6000, 700, 6, 256, "runs"),
])
def test_emulated_new_mpn_mul_product_kernels(emu, case, mode):
    """the pointwise products of new_mpn_mul (mul_fft.c:3244-3253) through the Karatsuba-block (2) and the
    schoolbook-block (3) kernel; all-ones operands make every half-sum a0+a1 / b0+b1 overflow"""
    n1, n2, depth, w, kind = case
    a, b = operand(kind, n1, 1), operand(kind, n2, 2)
    r = np.zeros(n1 + n2, dtype=np.uint64)
    emu.mpirfft_set_pointwise_mode(mode)
    try:
        emu.new_mpn_mul(ptr(r), ptr(a), n1, ptr(b), n2, depth, w)
    finally:
        emu.mpirfft_set_pointwise_mode(0)
    assert np.array_equal(r, L.gmp_mul(a, b))


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("l", [1, 3, 24, 64, 128, 256, 512])
def test_emulated_mulmod_adversarial(emu, l, mode):
    if mode >= 4 and l not in (256, 512):
        pytest.skip("two Karatsuba levels exist at 256 and 512 limbs")
    emu.mpirfft_set_pointwise_mode(mode)        # 0: default, 1: nested SS kernel, 2: Karatsuba blocks, 3: schoolbook blocks, 4 / 5: two Karatsuba levels
    random.seed(l)
    NW = 64 * l
    p = (1 << NW) + 1
    A = [random.getrandbits(NW) for _ in range(3)] + [p - 2, p - 2, p - 1, p - 1, 0, 1, p - 2, (1 << (NW // 2)) - 1, p - 2, (1 << NW) - (1 << (NW // 2))]
    B = [random.getrandbits(NW) for _ in range(3)] + [p - 2, 2, p - 1, 12345, 77, p - 2, 1 << (NW - 1), (1 << (NW // 2)) - 1, random.getrandbits(NW), (1 << NW) - (1 << (NW // 2))]
    a = np.stack([int_to_block(v, l) for v in A])
    b = np.stack([int_to_block(v, l) for v in B])
    da, db = emu.mpirfft_malloc_device(a.nbytes), emu.mpirfft_malloc_device(b.nbytes)
    emu.mpirfft_memcpy_h2d(da, ptr(a), a.nbytes, None)
    emu.mpirfft_memcpy_h2d(db, ptr(b), b.nbytes, None)
    assert emu.mpirfft_mulmod_batch_device(da, db, len(A), l, l + 1, None) == 0
    out = np.empty_like(a)
    emu.mpirfft_memcpy_d2h(ptr(out), da, a.nbytes, None)
    emu.mpirfft_set_pointwise_mode(0)
    for k in range(len(A)):
        assert np.array_equal(out[k], int_to_block(A[k] * B[k] % p, l)), (l, k, mode)


@pytest.mark.parametrize("l", [1024, 2048])
def test_emulated_mulmod_transform_path(emu, l):
    """residues of >= 250 limbs go through the batched negacyclic-transform path (csrc/host/mm.c,
    FFT_mulmod_2expp1 mul_fft.c:2998): random, all-ones, +-1 and 2^NW operands"""
    random.seed(l)
    NW = 64 * l
    p = (1 << NW) + 1
    A = [random.getrandbits(NW) for _ in range(2)] + [p - 1, p - 2, 0, 1, (1 << NW) - 1, p - 1, (1 << (NW // 2)) - 1, p - 2]
    B = [random.getrandbits(NW) for _ in range(2)] + [p - 2, p - 2, 5, p - 1, (1 << NW) - 1, p - 1, (1 << (NW // 2)) - 1, 2]
    a = np.stack([int_to_block(v, l) for v in A])
    b = np.stack([int_to_block(v, l) for v in B])
    da, db = emu.mpirfft_malloc_device(a.nbytes), emu.mpirfft_malloc_device(b.nbytes)
    emu.mpirfft_memcpy_h2d(da, ptr(a), a.nbytes, None)
    emu.mpirfft_memcpy_h2d(db, ptr(b), b.nbytes, None)
    assert emu.mpirfft_mulmod_batch_device(da, db, len(A), l, l + 1, None) == 0
    out = np.empty_like(a)
    emu.mpirfft_memcpy_d2h(ptr(out), da, a.nbytes, None)
    for k in range(len(A)):
        assert np.array_equal(out[k], int_to_block(A[k] * B[k] % p, l)), (l, k)


def test_emulated_mulmod_long_ripple(emu, monkeypatch):
    """all-ones operands with many small pieces (K = 128 chunks of 32 limbs): the recombination's
    carries cross almost the whole result, one chunk per round"""
    monkeypatch.setenv("MPIRFFT_MM_INNER", "64")
    l = 4096
    NW = 64 * l
    p = (1 << NW) + 1
    A = [p - 2, (1 << NW) - 1, (1 << (NW // 2)) - 1, p - 2, 12345]
    B = [p - 2, (1 << NW) - 1, (1 << (NW // 2)) - 1, 2, p - 2]
    a = np.stack([int_to_block(v, l) for v in A])
    b = np.stack([int_to_block(v, l) for v in B])
    da, db = emu.mpirfft_malloc_device(a.nbytes), emu.mpirfft_malloc_device(b.nbytes)
    emu.mpirfft_memcpy_h2d(da, ptr(a), a.nbytes, None)
    emu.mpirfft_memcpy_h2d(db, ptr(b), b.nbytes, None)
    assert emu.mpirfft_mulmod_batch_device(da, db, len(A), l, l + 1, None) == 0
    out = np.empty_like(a)
    emu.mpirfft_memcpy_d2h(ptr(out), da, a.nbytes, None)
    for k in range(len(A)):
        assert np.array_equal(out[k], int_to_block(A[k] * B[k] % p, l)), k


@pytest.mark.parametrize("n1,n2,kind", [(40000, 30000, "uniform"), (30000, 30000, "ones"), (9000, 3, "runs"), (150, 40, "uniform")])
def test_emulated_mpn_mul_wrapper(emu, n1, n2, kind):
    """mpirfft_mpn_mul chooses (depth, w) itself (smallest fused ring that is legal)"""
    a, b = operand(kind, n1, 11), operand(kind, n2, 12)
    r = np.zeros(n1 + n2, dtype=np.uint64)
    emu.mpirfft_mpn_mul(ptr(r), ptr(a), n1, ptr(b), n2)
    assert np.array_equal(r, L.gmp_mul(a, b))


def test_emulated_squaring_shortcut(emu):
    """device plan with d_i1 == d_i2: one forward transform"""
    n, depth, w = 30000, 10, 4
    a = operand("uniform", n, 21)
    h = C.c_void_p()
    assert emu.mpirfft_mul_plan_create(C.byref(h), n, n, depth, w) == 0
    da, dr = emu.mpirfft_malloc_device(a.nbytes), emu.mpirfft_malloc_device(2 * a.nbytes)
    emu.mpirfft_memcpy_h2d(da, ptr(a), a.nbytes, None)
    assert emu.mpirfft_mul_exec_device(h, dr, da, da, None) == 0
    r = np.zeros(2 * n, dtype=np.uint64)
    emu.mpirfft_memcpy_d2h(ptr(r), dr, r.nbytes, None)
    emu.mpirfft_mul_plan_destroy(h)
    assert np.array_equal(r, L.gmp_mul(a, a))


def test_emulated_result_table_rr_rs(emu):
    """FFT_radix2(rr, rs, ii, ...): results are delivered through rr with stride rs (mul_fft.c:786-827);
    with rr != ii the inputs' blocks keep their contents"""
    import ctypes as C
    from common import rand_blocks, cl, cul, residues
    n, w = 16, 8
    l, N = n * w // 64, 2 * n
    rng = np.random.default_rng(5)
    data = rand_blocks(rng, N, l)
    s1 = L.Slab(N, l, data)
    emu.FFT_radix2(s1.ii, cl(1), s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp)
    want = residues(s1.all(), l)
    s2, out = L.Slab(N, l, data), L.Slab(2 * N, l)
    emu.FFT_radix2(out.ii, cl(2), s2.ii, cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp)
    got = residues(out.all(), l)
    assert [got[2 * k] for k in range(N)] == want
    assert all(v == 0 for v in got[1::2])
    assert np.array_equal(s2.all(), data)


@pytest.mark.parametrize("inverse", [0, 1])
def test_emulated_sqrt2_mfa_odd_w_on_the_fused_path(emu, inverse):
    """FFT/IFFT_radix2_mfa_truncate_sqrt2 with odd w on a ring the fused tile executor serves (l = 64, w = 1):
    the even / odd column classes as two sub-batches of tile passes, against the compiled reference"""
    from common import rand_blocks, cl, cul, residues
    ref = L.load_ref(True)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    n, w, n1 = 4096, 1, 64
    trunc = 2 * n + 3 * 2 * n1
    l, N = n * w // 64, 4 * n
    n2, trunc2 = 2 * n // n1, (trunc - 2 * n) // n1
    depth = n2.bit_length() - 1
    rows = list(range(n2)) + [n2 + int(format(s, "0%db" % depth)[::-1], 2) for s in range(trunc2)]
    rng = np.random.default_rng(77 + inverse)
    data = rand_blocks(rng, N, l)
    if not inverse:
        data[trunc:] = 0
    name = ("I" if inverse else "") + "FFT_radix2_mfa_truncate_sqrt2"
    s1, s2 = L.Slab(N, l, data), L.Slab(N, l, data)
    getattr(ref, name)(s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp, cl(n1), cl(trunc))
    getattr(emu, name)(s2.ii, cl(n), cul(w), s2.pt1, s2.pt2, s2.ptmp, cl(n1), cl(trunc))
    idx = [i * n1 + j for i in rows for j in range(n1)] if not inverse else list(range(trunc))
    a1, a2 = s1.all(), s2.all()
    assert residues([a1[k] for k in idx], l) == residues([a2[k] for k in idx], l)


@pytest.mark.parametrize("total,bits,out", [(100, 29, 1), (300, 640, 12), (50, 64, 3)])
def test_emulated_combine_adds_to_res(emu, total, bits, out):
    """FFT_combine_bits / FFT_combine add the shifted coefficients to what res holds on entry
    (mul_fft.c:185, 229-233), modulo 2^(64 total_limbs)."""
    rng = np.random.default_rng(total)
    length = (64 * total - 1) // bits + 1
    s = L.Slab(length, out)
    coef = []
    for k in range(length):
        v = int(rng.integers(0, 2 ** 62)) % (1 << min(bits, 62))
        s.mem[k * (out + 1)] = v
        coef.append(v)
    r0 = rng.integers(0, 2 ** 64, total, dtype=np.uint64)
    r = r0.copy()
    cl = C.c_long
    emu.FFT_combine_bits.restype = None
    emu.FFT_combine_bits(ptr(r), s.ii, cl(length), cl(bits), cl(out), cl(total))
    want = (sum(int(x) << (64 * i) for i, x in enumerate(r0)) + sum(v << (bits * k) for k, v in enumerate(coef))) % (1 << (64 * total))
    got = sum(int(x) << (64 * i) for i, x in enumerate(r))
    assert got == want


@pytest.mark.parametrize("pinned", [False, True])
def test_emulated_host_copy_paths(pinned):
    """new_mpn_mul with operands in pageable memory goes through the worker threads and the pinned staging
    buffer (hostcopy.c; many small chunks here), with pinned operands straight to the device; both in a
    fresh process because the switches are read once."""
    code = r'''
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from oracle import loader as L
from common import operand, ptr
from mpir_fft_b200._lib import bind
emu = bind(C.CDLL(%r, mode=C.RTLD_LOCAL))
for n1, n2, depth, w in [(6000, 6000, 6, 256), (6000, 10, 6, 256), (3000, 3000, 7, 96), (2000, 1500, 6, 1024)]:
    for rep in range(2):
        a, b = operand("uniform", n1, 10 + rep), operand("runs", n2, 20 + rep)
        r = np.zeros(n1 + n2, dtype=np.uint64)
        emu.new_mpn_mul(ptr(r), ptr(a), n1, ptr(b), n2, depth, w)
        assert np.array_equal(r, L.gmp_mul(a, b)), (n1, n2, rep)
print("ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
       os.path.join(EMU_DIR, "libmpirfft_emu.so"))
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    env = dict(os.environ, MPIRFFT_COPY_THREADS="3", MPIRFFT_COPY_CHUNK="4096", MPIRFFT_COPY_MIN="0")
    if pinned:
        env["EMU_ALL_PINNED"] = "1"
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, (out.stdout[-500:], out.stderr[-800:])


def test_emulated_pointwise_tail_launch():
    """opt-in (MPIRFFT_PW_TAIL=1): the remainder of the last wave of the product kernel goes into a second launch
    of one-warp CTAs (mfft_dev_pointwise: PW_LAUNCH_K); the emulator pretends that eight warps fill the device"""
    code = r'''
import ctypes as C, os, random, sys
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from common import ptr, int_to_block
from mpir_fft_b200._lib import bind
emu = bind(C.CDLL(%r, mode=C.RTLD_LOCAL))
for l in (128, 256, 512):
    for count in (9, 10, 18):
        random.seed(l + count)
        NW = 64 * l; p = (1 << NW) + 1
        A = [random.getrandbits(NW) for _ in range(count - 2)] + [p - 1, p - 2]
        B = [random.getrandbits(NW) for _ in range(count - 2)] + [p - 2, p - 2]
        a = np.stack([int_to_block(v, l) for v in A]); b = np.stack([int_to_block(v, l) for v in B])
        da, db = emu.mpirfft_malloc_device(a.nbytes), emu.mpirfft_malloc_device(b.nbytes)
        emu.mpirfft_memcpy_h2d(da, ptr(a), a.nbytes, None); emu.mpirfft_memcpy_h2d(db, ptr(b), b.nbytes, None)
        emu.mpirfft_launch_count_reset()
        assert emu.mpirfft_mulmod_batch_device(da, db, count, l, l + 1, None) == 0
        assert emu.mpirfft_launch_count() == 2, emu.mpirfft_launch_count()
        out = np.empty_like(a); emu.mpirfft_memcpy_d2h(ptr(out), da, a.nbytes, None)
        for k in range(count):
            assert np.array_equal(out[k], int_to_block(A[k] * B[k] %% p, l)), (l, count, k)
print("ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
       os.path.join(EMU_DIR, "libmpirfft_emu.so"))
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MPIRFFT_PW_TAIL_TEST="1", MPIRFFT_PW_TAIL="1"),
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, (out.stdout[-500:], out.stderr[-800:])
