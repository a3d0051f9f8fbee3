import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """MPIRFFT_TEST_EMU=1 (developer aid, CPU only): run the gpu-marked parity tests against the
    CPU-emulated twin of the library (tests/emu) to debug the tests themselves.  Never set on the
    GPU box; the default is always the real library."""
    if os.environ.get("MPIRFFT_TEST_EMU") == "1":
        import ctypes
        import subprocess
        import mpir_fft_b200._lib as _l
        emu_dir = os.path.join(ROOT, "tests", "emu")
        subprocess.check_call(["make", "-s", "-C", emu_dir])
        _l._lib = _l.bind(ctypes.CDLL(os.path.join(emu_dir, "libmpirfft_emu.so"), mode=ctypes.RTLD_LOCAL))
