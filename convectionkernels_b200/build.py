"""Builds convectionkernels_b200/_build/libcvtt_b200.so (sm_100a only) with nvcc.

The flags are part of the numerical contract: -fmad=false (the reference's fp32 expressions
round after every operation), IEEE division and square root (nvcc defaults -prec-div=true -prec-sqrt=true, no
--use_fast_math), denormals kept (-ftz=false default).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libcvtt_b200.so")

SOURCES = ["cvtt_b200.cu", "bc7_host.cpp", "bc6h_host.cpp", "etc_host.cpp", "s3tc_host.cpp"]
HEADERS = ["cvtt_common.cuh", "bc7_core.cuh", "bc7_host.h", "bc7_tables.inc", "bc6h_core.cuh", "bc6h_host.h", "bc6h_tables.inc", "etc_core.cuh", "etc_host.h", "etc_tables.inc", "etc_bt709_table.inc", "s3tc_sc_tables.inc", "s3tc_core.cuh", "s3tc_host.h", os.path.join("..", "..", "include", "cvtt_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-msse2,-O2",
    "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compiles the library if it is missing or older than its sources.  Returns the path."""
    if not force and not is_stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libcvtt_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
