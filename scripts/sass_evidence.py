"""Regenerates profiles/rNN_sass_evidence.txt and profiles/rNN_ptxas_hot_kernels.log from the built library:
   python scripts/sass_evidence.py r02
(cuobjdump -sass instruction counts of the hot kernels; the ptxas -v lines of the same kernels from the build log)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "mpir_fft_b200", "libmpirfft_b200.so")
LOG = os.path.join(ROOT, "build", "obj", "mfft_kernels.ptxas.log")
HOT = [r"k_pointwiseILi16ELi1ELi4ELb0E", r"k_pointwiseILi32ELi1ELi4ELb0E", r"k_pointwiseILi8ELi1ELi4ELb0E",
       r"k_run_tilesILi4ELi256ELi2E", r"k_run_tiles_pILi4E", r"k_run_tiles_slicedILi1ELi256E", r"k_combine_sumadd",
       r"k_combine_sumPm", r"k_stage_cs_ip"]
KEYS = ["ACQBULK", "BAR", "IADD3", "IMAD.WIDE", "IMAD.WIDE.U32.X", "LDG", "LDGDEPBAR", "LDGSTS", "LDS", "SHFL", "STG", "STS",
        "SYNCS", "UBLKCP", "UTMALDG"]


def main(tag):
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    fn, per, whole = None, collections.OrderedDict(), collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            per[fn]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "IMAD.WIDE.U32.X" and op.startswith("IMAD.WIDE.U32.X")):
                    per[fn][k] += 1
                    whole[k] += 1
    out = ["# cuobjdump -sass mpir_fft_b200/libmpirfft_b200.so : instruction counts of the hot kernels (built from the committed sources; scripts/sass_evidence.py)",
           "# arch lines: " + ", ".join(arch)]
    for pat in HOT:
        for f, c in per.items():
            if pat in f:
                out.append(f)
                out.append("    total %d instructions; " % c["total"] + ", ".join("%s %d" % (k, c[k]) for k in KEYS if c[k]))
    out.append("whole library: " + ", ".join("%s %d" % (k, whole[k]) for k in ("ACQBULK", "IMAD.WIDE.U32.X", "LDGSTS", "SYNCS", "UBLKCP", "UTMALDG")))
    open(os.path.join(ROOT, "profiles", tag + "_sass_evidence.txt"), "w").write("\n".join(out) + "\n")
    lines = open(LOG).read().splitlines()
    keep = ["# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v  (csrc/Makefile), hot kernels only"]
    i = 0
    while i < len(lines):
        if "Compiling entry function" in lines[i] and any(p in lines[i] for p in HOT):
            j = i
            while j < len(lines) and (j == i or "Compiling entry function" not in lines[j]):
                keep.append(lines[j]); j += 1
            i = j
        else:
            i += 1
    open(os.path.join(ROOT, "profiles", tag + "_ptxas_hot_kernels.log"), "w").write("\n".join(keep) + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
