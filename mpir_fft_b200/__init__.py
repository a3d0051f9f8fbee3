"""mpir_fft_b200 -- B200-native (sm_100a) Schoenhage-Strassen multiplication behind the C entry
points of wbhart/mpir-fft.  The product is the C-ABI shared library libmpirfft_b200.so
(include/mpirfft_b200.h); this package is the thin Python mirror used by the tests and bench.py.
"""
from ._lib import lib, build, LIB_PATH, MulParams  # noqa: F401
from .api import (  # noqa: F401
    new_mpn_mul, new_mpn_mul6, mpn_mul, mul_params, choose_params, choose_params6, MulPlan, mulmod_batch, have_gpu, init,
)
