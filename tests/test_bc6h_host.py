"""CPU tests of the BC6H device logic: convectionkernels_b200/csrc/bc6h_core.cuh compiled for the CPU (tests/hostsim, test-only; the
eight lanes of a reference group run as eight threads and the kernel's segment ballots become a barrier vote) against the golden
vectors recorded from the unmodified reference."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, first_mismatch


@pytest.fixture(scope="module")
def hostsim_bc6h():
    out = os.path.join(ROOT, "tests", "_build", "libcvtt_hostsim_bc6h.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    csrc = os.path.join(ROOT, "convectionkernels_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off", "-msse2", "-pthread", "-w", "-I", csrc, "-o", out,
                           os.path.join(ROOT, "tests", "hostsim", "hostsim_bc6h.cpp"), os.path.join(csrc, "bc6h_host.cpp")])
    H = ctypes.CDLL(out)
    H.hostsim_encode_bc6h.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    H.hostsim_encode_bc6h_warp.argtypes = H.hostsim_encode_bc6h.argtypes
    H.hostsim_encode_bc6h_split.argtypes = H.hostsim_encode_bc6h.argtypes + [ctypes.c_int, ctypes.c_int]
    return H


@pytest.mark.parametrize("name", golden_names("bc6h"))
def test_bc6h_device_logic_on_cpu_matches_golden(hostsim_bc6h, name):
    g = load_golden(name)
    blocks = np.ascontiguousarray(g["blocks"])
    n = blocks.shape[0]
    out = np.zeros((n, 16), np.uint8)
    opt = np.ascontiguousarray(g["options"])
    rcp = np.ascontiguousarray(g["rcp"], dtype=np.float32)
    signed = 1 if str(g["fmt"]) == "BC6HS" else 0
    rc = hostsim_bc6h.hostsim_encode_bc6h(blocks.ctypes.data, n, out.ctypes.data, opt.ctypes.data, signed, rcp.ctypes.data)
    assert rc == 0
    assert (out == g["expected"]).all(), first_mismatch(g["expected"], out)


def test_integer_quantiser_equals_directed_rounding_fp32(hostsim_bc6h):
    """the kernel's ceil(N / 31) against the reference's fp32 multiply / divide / convert under MXCSR round-up, all inputs"""
    assert hostsim_bc6h.hostsim_bc6h_quantizer_mismatches() == 0


def test_half_conversion_reproduces_the_reference_formula(hostsim_bc6h):
    """TwosCLHalfToFloat is NOT the IEEE conversion for exponent 0 (it yields m * 2^-25); the kernels convert in hardware and
    halve those values"""
    assert hostsim_bc6h.hostsim_bc6h_half_conversion_mismatches() == 0


def test_random_blocks_with_denormals_and_wrapping_errors(hostsim_bc6h, reference):
    """Blocks of the GPU suite's random set that a plain IEEE half conversion gets wrong (half denormals) and that the exact
    pruning must not touch (signed fast indexing: SqDiffSInt16 yields negative error terms), simulated warp of four groups."""
    from convectionkernels_b200 import api, synth
    rcp = np.ascontiguousarray(reference.rcp_table(), dtype=np.float32)
    for fmt, signed, flags, picks in (("BC6HU", 0, None, (77, 203, 275, 491)), ("BC6HS", 1, None, (41, 401, 1216, 3262)),
                                      ("BC6HS", 1, api.Flags.Default | 0x240, (299, 587, 701, 749))):
        o = api.Options()
        if flags is not None:
            o.flags = flags
        opt = np.frombuffer(bytes(memoryview(o)), np.uint8).copy()
        all_blocks = synth.random_blocks_f16(4096 + 8, seed=77, signed=bool(signed))
        idx = (np.array([b // 8 for b in picks])[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        blocks = np.ascontiguousarray(all_blocks[idx])
        want = reference.encode(fmt, blocks, opt)
        out = np.zeros_like(want)
        assert hostsim_bc6h.hostsim_encode_bc6h_warp(blocks.ctypes.data, len(blocks), out.ctypes.data, opt.ctypes.data, signed, rcp.ctypes.data) == 0
        assert (out == want).all(), (fmt, flags, first_mismatch(want, out))


def test_image_content_against_reference(hostsim_bc6h, reference):
    """The device code (compiled for the CPU) against the unmodified reference running in this process on an HDR ramp crop and
    on random halves, unsigned and signed: exercises the pruned commit scan (rows that cannot beat the lane's best are
    skipped) on data the fixtures do not contain.  The rcp table is this host's, like the reference's own _mm_rcp_ps."""
    from convectionkernels_b200 import api, synth
    opt = np.frombuffer(bytes(memoryview(api.Options())), np.uint8).copy()
    rcp = np.ascontiguousarray(reference.rcp_table(), dtype=np.float32)      # this host's _mm_rcp_ps, as the reference itself uses it
    for fmt, signed in (("BC6HU", 0), ("BC6HS", 1)):
        img = synth.image_to_blocks(synth.hdr_ramp_f16(96, 96, seed=99, signed=bool(signed)))
        rnd = synth.image_to_blocks(synth.hdr_ramp_f16(64, 64, seed=3, signed=bool(signed)))[::-1]
        blocks = np.ascontiguousarray(np.concatenate([img, rnd]))
        blocks = blocks[: len(blocks) // 8 * 8]
        want = reference.encode(fmt, blocks, opt)
        out = np.zeros_like(want)
        assert hostsim_bc6h.hostsim_encode_bc6h(blocks.ctypes.data, len(blocks), out.ctypes.data, opt.ctypes.data, signed, rcp.ctypes.data) == 0
        assert (out == want).all(), (fmt, first_mismatch(want, out))


def test_pruning_does_not_change_results(hostsim_bc6h, monkeypatch):
    """bc6h_partition skips work that provably cannot change the result (P.prune); with the skipping switched off
    (CVTTB200_BC6H_NO_PRUNE=1, the A/B timing knob) the bytes must be the same."""
    g = load_golden("bc6hu_random")
    blocks = np.ascontiguousarray(g["blocks"])
    n = blocks.shape[0]
    opt = np.ascontiguousarray(g["options"])
    rcp = np.ascontiguousarray(g["rcp"], dtype=np.float32)
    monkeypatch.setenv("CVTTB200_BC6H_NO_PRUNE", "1")
    out = np.zeros((n, 16), np.uint8)
    assert hostsim_bc6h.hostsim_encode_bc6h(blocks.ctypes.data, n, out.ctypes.data, opt.ctypes.data, 0, rcp.ctypes.data) == 0
    assert (out == g["expected"]).all(), first_mismatch(g["expected"], out)


@pytest.mark.parametrize("name,calls_per_slice,seed_calls", [(n, c, k) for n in golden_names("bc6h") for c, k in ((4, 4), (33, 0))] + [("bc6hu_random", 1, 0), ("bc6hs_random", 196, 0), ("bc6hs_random", 7, 4), ("bc6hu_random", 2, 4)])
def test_small_call_launch_logic_matches_golden(hostsim_bc6h, name, calls_per_slice, seed_calls):
    """The small-call launch (bc6h_kernels.cu): the 196 calls of the search in independent ranges that only record error
    histories, then one re-run of each lane's winner call for its whole group with the true entry errors (bc6h_core.cuh,
    "The search as a numbered sequence of CALLS"), with and without the ranges starting from the error of the four one-subset
    calls.  Same bytes as the sequential search, group coupling included."""
    if name not in golden_names("bc6h"):
        pytest.skip("no such fixture")
    g = load_golden(name)
    blocks = np.ascontiguousarray(g["blocks"])
    n = blocks.shape[0]
    out = np.zeros((n, 16), np.uint8)
    opt = np.ascontiguousarray(g["options"])
    rcp = np.ascontiguousarray(g["rcp"], dtype=np.float32)
    signed = 1 if str(g["fmt"]) == "BC6HS" else 0
    assert hostsim_bc6h.hostsim_encode_bc6h_split(blocks.ctypes.data, n, out.ctypes.data, opt.ctypes.data, signed, rcp.ctypes.data, calls_per_slice, seed_calls) == 0
    assert (out == g["expected"]).all(), first_mismatch(g["expected"], out)
