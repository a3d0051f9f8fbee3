"""CPU: the host-side parts of bench.py -- the reference arm's JSON line (the compiled reference or
the oracle port timed on the host cores, no GPU involved) and the clock sampler's behaviour on a
machine without NVML / nvidia-smi (it must report zero samples, not hang or raise)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [x for x in out.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mlimb/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "65536 x 65536" in d["config"]["workload"]


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--workload", "cfg1", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_without_gpu():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    t = time.time()
    s.begin()
    c = s.finish()
    assert time.time() - t < 10
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
    if c["samples"] == 0:
        assert c["sm_mhz"] is None and c["reasons"] == []
