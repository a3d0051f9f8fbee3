"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol that
include/mpirfft_b200.h declares, validates parameters, and has no CPU compute path."""
import os
import re
import subprocess

import numpy as np
import pytest

import mpir_fft_b200 as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mpirfft_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    skip = {"defined", "if", "sizeof"}
    return sorted({n for n in names if n not in skip and not n.isupper()})


def test_every_declared_symbol_is_exported():
    lib = M.lib()
    syms = declared_symbols()
    assert len(syms) >= 50
    for s in ["new_mpn_mul", "new_mpn_mulmod_2expp1", "fft_mulmod_2expp1", "FFT_mulmod_2expp1",
              "FFT_radix2_mfa_truncate", "IFFT_radix2_mfa_truncate", "FFT_radix2_truncate_twiddle",
              "IFFT_radix2_truncate1_twiddle", "FFT_radix2_butterfly", "FFT_split_bits", "FFT_combine_bits"]:
        assert s in syms
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, "declared in include/mpirfft_b200.h but not exported: %s" % missing


def test_library_is_sm100a_native():
    out = subprocess.run(["cuobjdump", "--list-elf", M.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(5|6|7|8|9)\d", out), "only sm_100a code is built (no multi-arch fallback)"


def test_parameter_derivation_matches_reference_table():
    # SURVEY section 8 table (mul_fft.c:3193-3203)
    p = M.mul_params(1 << 16, 1 << 16, 12, 1)
    assert (p["bits1"], p["j1"], p["trunc"], p["limbs"], p["sqrt"], p["n2"], p["trunc_rows"]) == (2042, 2055, 4224, 64, 64, 128, 66)
    p = M.mul_params(1 << 20, 1 << 20, 14, 1)
    assert (p["bits1"], p["j1"], p["trunc"], p["limbs"], p["trunc_rows"]) == (8185, 8200, 16640, 256, 130)
    p = M.mul_params(3000000, 1700000, 14, 2)
    assert (p["bits1"], p["j1"], p["j2"], p["trunc"], p["limbs"]) == (16377, 11724, 6644, 18432, 512)
    p = M.mul_params(3000000, 1700000, 15, 1)
    assert (p["bits1"], p["j1"], p["j2"], p["trunc"], p["limbs"]) == (16376, 11725, 6644, 18432, 512)
    p = M.mul_params(1 << 30, 1 << 30, 19, 1)
    assert (p["bits1"], p["j1"], p["trunc"], p["limbs"], p["sqrt"], p["n2"], p["trunc_rows"]) == (262134, 262155, 525312, 8192, 512, 2048, 1026)


def test_illegal_parameters():
    with pytest.raises(ValueError):
        M.mul_params(1 << 17, 1 << 17, 11, 4)      # 4103 > 4096 coefficients: the reference segfaults
    with pytest.raises(ValueError):
        M.mul_params(8, 8, 5, 1)                   # 64 does not divide n*w
    with pytest.raises(ValueError):
        M.mul_params(0, 8, 6, 1)
    assert M.choose_params(1 << 20, 1 << 20) == (14, 1)
    assert M.choose_params(3000000, 1700000) == (14, 2)


@pytest.mark.skipif(M.have_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    a = np.ones(20, dtype=np.uint64)
    with pytest.raises(RuntimeError):
        M.new_mpn_mul(a, a, 6, 1)
    with pytest.raises(RuntimeError):
        M.MulPlan(20, 20, 6, 1)
    # the raw C symbol aborts the process with a diagnostic
    code = ("import numpy as np, ctypes as C, mpir_fft_b200 as M; a=np.ones(20,dtype=np.uint64); r=np.zeros(40,dtype=np.uint64);"
            "M.lib().new_mpn_mul(r.ctypes.data,a.ctypes.data,20,a.ctypes.data,20,6,1)")
    pr = subprocess.run(["python", "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert pr.returncode != 0 and "no usable CUDA device" in pr.stderr


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mpir_fft_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt and "libgmp" not in txt, os.path.join(dirpath, f)
