"""CPU tests of the BC1-BC5 device logic: convectionkernels_b200/csrc/s3tc_core.cuh compiled for the CPU (tests/hostsim, test-only)
against the golden vectors recorded from the unmodified reference (incl. BASELINE.json configs[0]: EncodeBC1 on the 256x256
gradient)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, first_mismatch
from oracle.loader import FMT


@pytest.fixture(scope="module")
def hostsim_s3tc():
    out = os.path.join(ROOT, "tests", "_build", "libcvtt_hostsim_s3tc.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    csrc = os.path.join(ROOT, "convectionkernels_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off", "-msse2", "-pthread", "-w", "-DCVTT_HOSTSIM", "-I", csrc, "-o", out,
                           os.path.join(ROOT, "tests", "hostsim", "hostsim_s3tc.cpp"), os.path.join(csrc, "s3tc_host.cpp")])
    H = ctypes.CDLL(out)
    H.hostsim_encode_s3tc.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return H


@pytest.mark.parametrize("name", [n for p in ("bc1_", "bc2_", "bc3_", "bc4", "bc5") for n in golden_names(p)])
def test_s3tc_device_logic_on_cpu_matches_golden(hostsim_s3tc, name):
    g = load_golden(name)
    blocks = np.ascontiguousarray(g["blocks"])
    n = blocks.shape[0]
    out = np.zeros_like(g["expected"])
    opt = np.ascontiguousarray(g["options"])
    rcp = np.ascontiguousarray(g["rcp"], dtype=np.float32)
    rc = hostsim_s3tc.hostsim_encode_s3tc(FMT[str(g["fmt"])], blocks.ctypes.data, n, out.ctypes.data, opt.ctypes.data, rcp.ctypes.data)
    assert rc == 0
    assert (out == g["expected"]).all(), first_mismatch(g["expected"], out)
