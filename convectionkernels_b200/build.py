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

SOURCES = ["bc7_kernels.cu", "etc_kernels.cu", "bc6h_kernels.cu", "s3tc_kernels.cu", "cvtt_b200.cu", "bc7_host.cpp", "bc6h_host.cpp", "etc_host.cpp", "s3tc_host.cpp"]
HEADERS = ["cvtt_common.cuh", "cvtt_internal.h", "cvtt_segment.cuh", "decode_core.cuh", "bc7_core.cuh", "bc7_host.h", "bc7_tables.inc", "bc6h_core.cuh", "bc6h_host.h", "bc6h_tables.inc", "etc_core.cuh", "etc_host.h", "etc_tables.inc", "etc_bt709_table.inc", "s3tc_sc_tables.inc", "s3tc_core.cuh", "s3tc_host.h", os.path.join("..", "..", "include", "cvtt_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-msse2,-O2",
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
    # one object per translation unit, compiled in parallel (the kernel TUs are independent: no relocatable device code),
    # objects whose sources and headers are unchanged are reused
    header_time = max(os.path.getmtime(p) for p in [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)] if os.path.exists(p))
    jobs = []
    objects = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, os.path.splitext(src)[0] + ".o")
        objects.append(obj)
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), header_time):
            continue
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
        jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, proc in jobs:
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out)
        if proc.returncode != 0:
            failed.append(src)
    if failed:
        raise RuntimeError("nvcc failed compiling " + ", ".join(failed))
    res = subprocess.run([nvcc_path(), "-shared", "-o", LIB] + objects, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed linking libcvtt_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
