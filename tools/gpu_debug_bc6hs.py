#!/usr/bin/env python3
"""GPU-side diagnosis of BC6H mismatches on the random blocks of tests/test_bc6h_gpu.py: which blocks differ, with and
without the exact pruning, alone and inside their warp / CTA."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from convectionkernels_b200 import api, synth, build as _b
if os.environ.get("CVTT_LIB_VARIANT"):
    _v = os.path.join(_b.OUT_DIR, os.environ["CVTT_LIB_VARIANT"])
    _b.LIB = _v
    _b.build = lambda *a, **k: _v
    print("library:", _v)
from oracle.loader import Reference
R = Reference()
api.init(0)
small = len(sys.argv) > 1 and sys.argv[1] == "small"
res = {}
for fmt, flags in (("BC6HU", None), ("BC6HS", None), ("BC6HS", api.Flags.Default | 0x240), ("BC6HU", api.Flags.Default | 0x40)):
    blocks = synth.random_blocks_f16(4096 + 8, seed=77, signed=fmt.endswith("S"))
    o = api.Options()
    if flags is not None:
        o.flags = flags
    ob = np.frombuffer(bytes(memoryview(o)), np.uint8)
    if small:
        got = api.encode(fmt, blocks[256:384], o)
        print(fmt, flags, "CTA 2 alone done")
        continue
    want = R.encode(fmt, blocks, ob, threads=0)
    for prune in ("0", "1"):
        os.environ["CVTTB200_BC6H_NO_PRUNE"] = prune
        got = api.encode(fmt, blocks, o)
        bad = np.nonzero((got != want).any(axis=1))[0]
        print(fmt, flags, "NO_PRUNE=" + prune, "differing:", bad.tolist())
        res["%s_%s_%s" % (fmt, flags, prune)] = got
    os.environ["CVTTB200_BC6H_NO_PRUNE"] = "0"
    got = api.encode(fmt, blocks, o)
    bad = np.nonzero((got != want).any(axis=1))[0]
    for b in ([] if os.environ.get("CVTT_LIB_VARIANT") else bad[:3]):
        g0, w0, c0 = (b // 8) * 8, (b // 32) * 32, (b // 128) * 128
        for name, lo, n in (("group", g0, 8), ("warp", w0, 32), ("cta", c0, 128)):
            alone = api.encode(fmt, blocks[lo:lo + n], o)
            print("   block", b, name, "alone: differing", np.nonzero((alone != want[lo:lo + n]).any(axis=1))[0].tolist())
        again = api.encode(fmt, blocks, o)
        print("   deterministic:", bool((again == got).all()))
    res["%s_%s_want" % (fmt, flags)] = want
if not small:
    np.savez_compressed("gpurun_out/debug_bc6h_random.npz", **res)
