"""The C++ source-compatibility shim (include/cvtt_b200_dropin.h): a reference-style caller compiles and links against
libcvtt_b200.so; without a GPU it aborts loudly (no CPU fallback), with one its output equals the oracle's."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from convectionkernels_b200 import api


def _build():
    api._lib()
    out = os.path.join(ROOT, "tests", "_build", "dropin_main")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    libdir = os.path.dirname(api.library_path())
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"),
                           "-o", out, "-L", libdir, "-lcvtt_b200", "-Wl,-rpath," + libdir, "-pthread"])
    return out


def _image_blocks(w, h):
    y, x = np.mgrid[0:h, 0:w]
    a = np.where(((x // 32) & 1) == 1, 255, 255 - y)
    img = np.stack([x & 255, y & 255, ((x + y) // 2) & 255, a & 255], axis=-1).astype(np.uint8)
    return img.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 4)


def _fnv64(data):
    h = 1469598103934665603
    for b in data.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful without a GPU")
def test_dropin_compiles_and_fails_loudly_without_gpu():
    exe = _build()
    r = subprocess.run([exe, "32", "32", "10"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode != 0
    assert "status -3" in r.stderr


@pytest.mark.gpu
def test_dropin_matches_oracle(oracle):
    exe = _build()
    r = subprocess.run([exe, "128", "64", "60"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    blocks = _image_blocks(128, 64)
    opt = np.frombuffer(bytes(memoryview(api.Options())), np.uint8)
    want = oracle.encode_bc7(blocks, opt, oracle.plan_from_quality(60))
    assert "%d blocks, fnv64 %016x" % (len(blocks), _fnv64(want)) in r.stdout, r.stdout
    from oracle import loader
    if os.path.exists(loader.REF_SO):
        R = loader.Reference()
        ref = R.encode("BC3", blocks, opt)
        assert "BC3 fnv64 %016x" % _fnv64(ref) in r.stdout, r.stdout
        enc, alloc = api.Options(), api.Options()
        enc.redWeight, enc.greenWeight, enc.blueWeight = 1.0, 0.5, 0.25
        alloc.redWeight, alloc.greenWeight, alloc.blueWeight = 0.1, 1.0, 0.7
        ref = R.encode("ETC2", blocks, np.frombuffer(bytes(memoryview(enc)), np.uint8), etc2_alloc_options=np.frombuffer(bytes(memoryview(alloc)), np.uint8))
        assert "ETC2 alloc-options fnv64 %016x" % _fnv64(ref) in r.stdout, r.stdout
    assert "4 caller threads OK" in r.stdout, r.stdout
