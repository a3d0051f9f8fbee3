"""CPU: the worker pool behind the pageable-memory path of new_mpn_mul (csrc/host/hostcopy.c): tens of
thousands of small jobs back to back, every chunk must run exactly once (a thread that is late leaving one
job must neither claim nor skip a chunk of the next)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
HOST = os.path.join(os.path.dirname(HERE), "mpir_fft_b200", "csrc", "host")


def test_worker_pool_runs_every_chunk_exactly_once(tmp_path):
    src = open(os.path.join(HOST, "hostcopy.c")).read().replace('#include "runtime.h"', "int mfft_dev_bind(void);")
    (tmp_path / "hc.c").write_text(src)
    exe = str(tmp_path / "stress")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-I", HOST, os.path.join(HERE, "hostcopy_stress.c"),
                           str(tmp_path / "hc.c"), "-o", exe, "-lpthread"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), (out.stdout[-300:], out.stderr[-300:])
