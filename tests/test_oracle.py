"""CPU: pin the checkers.  (1) the compiled reference passes the reference's own unit tests and,
once mul_fft.c:3246 is fixed, equals GMP mpn_mul while the raw one does not;
(2) the plain-C restatement oracle/ssmul_oracle.c equals the compiled reference function by
function, GMP for products, and the golden fixtures."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from oracle import loader as L
from common import rand_blocks, residues, operand, ptr, cl, cul, block_to_int, int_to_block
from golden import make_golden as G

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.npz"))
KIND = {"FFT_radix2": 0, "FFT_radix2_truncate": 1, "FFT_radix2_truncate1": 2, "IFFT_radix2": 3,
        "IFFT_radix2_truncate": 4, "IFFT_radix2_truncate1": 5}


@pytest.fixture(scope="module")
def ref():
    r = L.load_ref(True)
    if r is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return r


@pytest.fixture(scope="module")
def port():
    p = L.load_port()
    p.orc_new_mpn_mul.restype = C.c_int
    return p


@pytest.mark.parametrize("name", ["test_norm", "test_lshB_sumdiffmod", "test_fft_ifft", "test_fft_ifft_mfa"])
def test_reference_self_tests_pass(ref, name):
    getattr(ref, name)()          # abort()s on failure


def test_reference_bug_3246(ref):
    raw = L.load_ref(False)
    n = 1 << 12
    a, b = operand("uniform", n, 1), operand("uniform", n, 2)
    want = L.gmp_mul(a, b)
    r = np.zeros(2 * n, dtype=np.uint64)
    ref.new_mpn_mul(ptr(r), ptr(a), cl(n), ptr(b), cl(n), cul(10), cul(1))
    assert np.array_equal(r, want)
    r2 = np.zeros(2 * n, dtype=np.uint64)
    raw.new_mpn_mul(ptr(r2), ptr(a), cl(n), ptr(b), cl(n), cul(10), cul(1))
    assert not np.array_equal(r2, want), "the unpatched reference is expected to be wrong (revbin width)"


def _canon(blocks, l):
    return np.stack([int_to_block(block_to_int(b, l), l) for b in blocks])


@pytest.mark.parametrize("case", G.TRANSFORMS)
def test_port_transforms_vs_golden(port, case):
    name, n, w, trunc, seed = case
    if "negacyclic" in name:
        pytest.skip("negacyclic is checked through the schedule tests")
    data, l = G.transform_inputs(name, n, w, trunc, seed)
    s = L.Slab(2 * n, l, data)
    port.orc_transform(KIND[name], s.ii, cl(1), cl(n), cul(w), cul(0), cul(0), cul(0), cul(0), cl(trunc))
    upto = trunc if trunc else 2 * n
    assert np.array_equal(_canon(s.all()[:upto], l), GOLD["T_%s_%d_%d_%d" % (name, n, w, trunc)])


@pytest.mark.parametrize("case", G.MFAS)
def test_port_mfa_vs_golden(port, case):
    inverse, n, w, n1, trunc, seed = case
    data, l = G.mfa_inputs(inverse, n, w, n1, trunc, seed)
    s = L.Slab(2 * n, l, data)
    port.orc_mfa(inverse, s.ii, cl(n), cul(w), cl(n1), cl(trunc))
    gold = GOLD["M_%d_%d_%d_%d_%d" % (inverse, n, w, n1, trunc)]
    n2 = 2 * n // n1
    d = n2.bit_length() - 1
    rows = [int(format(x, "0%db" % d)[::-1], 2) for x in range(trunc // n1 if trunc else n2)]
    idx = [i * n1 + j for i in rows for j in range(n1)] if not inverse else list(range(trunc if trunc else 2 * n))
    got = _canon(s.all(), l)
    assert np.array_equal(got[idx], gold[idx])


@pytest.mark.parametrize("case", G.PRODUCTS)
def test_port_products_vs_golden_and_gmp(port, case):
    n1, n2, depth, w, kind = case
    a, b = operand(kind, n1, 0x5EED0001), operand(kind, n2, 0x5EED0002)
    r = np.zeros(n1 + n2, dtype=np.uint64)
    assert port.orc_new_mpn_mul(ptr(r), ptr(a), cl(n1), ptr(b), cl(n2), cul(depth), cul(w)) == 0
    assert np.array_equal(r, L.gmp_mul(a, b))
    key = "P_%d_%d_%d_%d_%s" % (n1, n2, depth, w, kind)
    if key in GOLD.files:
        assert np.array_equal(r, GOLD[key])
    else:
        assert hashlib.sha256(r.tobytes()).digest() == GOLD[key + "_sha256"].tobytes()


def test_port_vs_compiled_reference_live(ref, port):
    rng = np.random.default_rng(77)
    for name, k in KIND.items():
        for (n, w, trunc) in [(16, 4, 18), (64, 3, 100), (32, 64, 40)]:
            l, N = n * w // 64, 2 * n
            data = rand_blocks(rng, N, l)
            if name == "FFT_radix2_truncate":
                data[trunc:] = 0
            s1, s2 = L.Slab(N, l, data), L.Slab(N, l, data)
            args = [s1.ii, cl(1), s1.ii, cl(n), cul(w), s1.pt1, s1.pt2, s1.ptmp] + ([cl(trunc)] if "trunc" in name else [])
            getattr(ref, name)(*args)
            port.orc_transform(k, s2.ii, cl(1), cl(n), cul(w), cul(0), cul(0), cul(0), cul(0), cl(trunc))
            upto = trunc if "trunc" in name else N
            assert residues(s1.all()[:upto], l) == residues(s2.all()[:upto], l), (name, n, w, trunc)


def test_port_rejects_illegal_parameters(port):
    a = np.ones(1 << 17, np.uint64)
    r = np.zeros(1 << 18, np.uint64)
    assert port.orc_new_mpn_mul(ptr(r), ptr(a), cl(1 << 17), ptr(a), cl(1 << 17), cul(11), cul(4)) == -1
