"""CPU tests of the decoders' per-thread device code (convectionkernels_b200/csrc/decode_core.cuh compiled for the CPU, tests/hostsim;
test-only) against the unmodified reference's DecodeBC7 / DecodeBC6HU / DecodeBC6HS."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden

KIND = {"BC7": 0, "BC6HU": 1, "BC6HS": 2}


@pytest.fixture(scope="module")
def hostsim_decode():
    out = os.path.join(ROOT, "tests", "_build", "libcvtt_hostsim_decode.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    csrc = os.path.join(ROOT, "convectionkernels_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off", "-msse2", "-w", "-DCVTT_HOSTSIM", "-I", csrc, "-o", out,
                           os.path.join(ROOT, "tests", "hostsim", "hostsim_decode.cpp"), os.path.join(csrc, "bc7_host.cpp"), os.path.join(csrc, "bc6h_host.cpp")])
    H = ctypes.CDLL(out)
    H.hostsim_decode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    return H


def random_encoded_blocks(fmt, n, seed):
    """uniformly random 128-bit blocks; for BC7 the mode bit is forced so that all eight modes and the invalid mode byte are hit equally"""
    rng = np.random.default_rng(seed)
    bc = rng.integers(0, 256, size=(n, 16), dtype=np.uint8)
    if fmt == "BC7":
        mode = rng.integers(0, 9, size=n)
        low = ((1 << (mode + 1)) - 1) & 0xff
        bc[:, 0] = (bc[:, 0] & ~low & 0xff) | ((1 << mode) & 0xff)
    return bc


def _decode(H, fmt, bc):
    bc = np.ascontiguousarray(bc).reshape(-1, 16)
    out = np.zeros((bc.shape[0], 16, 4), np.uint8 if fmt == "BC7" else np.int16)
    assert H.hostsim_decode(KIND[fmt], bc.ctypes.data, bc.shape[0], out.ctypes.data) == 0
    return out


@pytest.mark.parametrize("fmt", ["BC7", "BC6HU", "BC6HS"])
def test_random_bit_patterns_decode_like_the_reference(hostsim_decode, reference, fmt):
    bc = random_encoded_blocks(fmt, 32768, seed=7)
    want = reference.decode(fmt, bc)
    got = _decode(hostsim_decode, fmt, bc)
    assert (got == want).all()


@pytest.mark.parametrize("name", golden_names("bc7_") + golden_names("bc6h"))
def test_golden_encodings_decode_like_the_reference(hostsim_decode, reference, name):
    g = load_golden(name)
    fmt = str(g["fmt"])
    want = reference.decode(fmt, g["expected"])
    got = _decode(hostsim_decode, fmt, g["expected"])
    assert (got == want).all()


def test_special_blocks(hostsim_decode, reference):
    """all-zero block (no BC7 mode bit), all-ones, reserved BC6H mode ids 0x13 0x17 0x1b 0x1f"""
    bc = np.zeros((8, 16), np.uint8)
    bc[1] = 0xff
    for i, m in enumerate((0x13, 0x17, 0x1b, 0x1f)):
        bc[2 + i] = 0xa5
        bc[2 + i, 0] = 0xe0 | m
    for fmt in ("BC7", "BC6HU", "BC6HS"):
        assert (_decode(hostsim_decode, fmt, bc) == reference.decode(fmt, bc)).all()
    assert (_decode(hostsim_decode, "BC7", bc)[0] == 0).all()
    assert (_decode(hostsim_decode, "BC6HU", bc)[2] == np.array([0, 0, 0, 0x3c00], np.int16)).all()
