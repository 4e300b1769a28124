"""CPU tests of the ETC device logic: convectionkernels_b200/csrc/etc_core.cuh compiled for the CPU (tests/hostsim, test-only; eight
threads per reference group, the kernel's segment maximum becomes a barrier reduction) against the golden vectors recorded from the
unmodified reference."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, first_mismatch

KIND = {"ETC1": 0, "ETC2": 1, "ETC2_RGBA": 2, "ETC2_ALPHA": 3, "EAC_R11U": 4, "EAC_R11S": 5, "ETC2_PUNCHTHROUGH": 6}


@pytest.fixture(scope="module")
def hostsim_etc():
    out = os.path.join(ROOT, "tests", "_build", "libcvtt_hostsim_etc.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    csrc = os.path.join(ROOT, "convectionkernels_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off", "-msse2", "-pthread", "-w", "-DCVTT_HOSTSIM", "-I", csrc, "-o", out,
                           os.path.join(ROOT, "tests", "hostsim", "hostsim_etc.cpp"), os.path.join(csrc, "etc_host.cpp")])
    H = ctypes.CDLL(out)
    H.hostsim_encode_etc.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    H.hostsim_encode_etc_alloc.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return H


@pytest.mark.parametrize("name", golden_names("etc") + golden_names("eac"))
def test_etc_device_logic_on_cpu_matches_golden(hostsim_etc, name):
    g = load_golden(name)
    blocks = np.ascontiguousarray(g["blocks"])
    n = blocks.shape[0]
    out = np.zeros_like(g["expected"])
    opt = np.ascontiguousarray(g["options"])
    rc = hostsim_etc.hostsim_encode_etc(KIND[str(g["fmt"])], blocks.ctypes.data, n, out.ctypes.data, opt.ctypes.data)
    assert rc == 0
    assert (out == g["expected"]).all(), first_mismatch(g["expected"], out)


def test_alloc_time_options_fix_the_chroma_axes(hostsim_etc, reference):
    """AllocETC2Data(options) derives the chroma side axes of the T / H search (ETC.cpp:3117-3145); EncodeETC2 takes flags and
    error weights from its own options (ETC.cpp:1773).  A caller that passes different weights to the two must get the
    reference's bytes -- which differ from both 'same options' encodings."""
    from convectionkernels_b200 import api, synth
    blocks = synth.random_blocks_rgba8(256, seed=41)
    enc, alloc = api.Options(), api.Options()
    enc.redWeight, enc.greenWeight, enc.blueWeight = 1.0, 0.5, 0.25
    alloc.redWeight, alloc.greenWeight, alloc.blueWeight = 0.1, 1.0, 0.7
    eb, ab = (np.frombuffer(bytes(memoryview(o)), np.uint8).copy() for o in (enc, alloc))
    want = reference.encode("ETC2", blocks, eb, etc2_alloc_options=ab)
    assert (want != reference.encode("ETC2", blocks, eb)).any() and (want != reference.encode("ETC2", blocks, ab)).any()
    out = np.zeros_like(want)
    assert hostsim_etc.hostsim_encode_etc_alloc(1, blocks.ctypes.data, len(blocks), out.ctypes.data, eb.ctypes.data, ab.ctypes.data) == 0
    assert (out == want).all(), first_mismatch(want, out)


def test_t_mode_group_coupling(hostsim_etc, reference):
    """the T-mode candidate list depends on the other blocks of the group (never-written slot = colour 0)"""
    from convectionkernels_b200 import api, synth
    base = synth.random_blocks_rgba8(64, seed=23)
    flat = np.broadcast_to(np.array([200, 30, 30, 255], np.uint8), (1, 16, 4)).copy()
    blocks = np.concatenate([np.concatenate([flat if k % 2 else base[k:k + 1], base[8 * (k % 8):8 * (k % 8) + 7]]) for k in range(16)])
    opt = np.frombuffer(bytes(memoryview(api.Options())), np.uint8)
    want = reference.encode("ETC2", blocks, opt)
    out = np.zeros_like(want)
    assert hostsim_etc.hostsim_encode_etc(1, blocks.ctypes.data, len(blocks), out.ctypes.data, opt.ctypes.data) == 0
    assert (out == want).all(), first_mismatch(want, out)


@pytest.mark.parametrize("flags,threshold", [(0x108, 0.5), (0x308, -1.0), (0xD08, 1.0)])
def test_punchthrough_group_coupling(hostsim_etc, reference, flags, threshold):
    """EncodeETC2PunchthroughAlpha: which stages run depends on the other blocks of the group (ETC.cpp:1720,1850,1865), and the
    virtual T mode walks offsets up to the group's largest line-pixel count (ETC.cpp:1017-1044)"""
    from convectionkernels_b200 import api, synth
    blocks = synth.punchthrough_blocks_rgba8(1024, seed=77)
    o = api.Options()
    o.flags, o.threshold = flags, threshold
    opt = np.frombuffer(bytes(memoryview(o)), np.uint8)
    want = reference.encode("ETC2_PUNCHTHROUGH", blocks, opt)
    out = np.zeros_like(want)
    assert hostsim_etc.hostsim_encode_etc(6, blocks.ctypes.data, len(blocks), out.ctypes.data, opt.ctypes.data) == 0
    assert (out == want).all(), first_mismatch(want, out)


def test_image_content_against_reference(hostsim_etc, reference):
    """The device code (compiled for the CPU) against the unmodified reference on image-like and random content: covers the
    bookkeeping the golden fixtures are too small to stress (differential attempts filtered while generated, the H-mode floor
    test that skips hopeless pair loops, multiply-high divisions, pairwise error evaluation)."""
    from convectionkernels_b200 import api, synth
    opt = np.frombuffer(bytes(memoryview(api.Options())), np.uint8).copy()
    blocks = np.ascontiguousarray(np.concatenate([synth.image_to_blocks(synth.mixed_rgba8(128, 128, seed=1234)), synth.random_blocks_rgba8(512, seed=9)]))
    for fmt, kind in (("ETC2", 1), ("ETC2_RGBA", 2), ("ETC1", 0)):
        want = reference.encode(fmt, blocks, opt)
        out = np.zeros_like(want)
        assert hostsim_etc.hostsim_encode_etc(kind, blocks.ctypes.data, len(blocks), out.ctypes.data, opt.ctypes.data) == 0
        assert (out == want).all(), (fmt, first_mismatch(want, out))
