#!/usr/bin/env python3
"""Dev check (GPU box): BC7 kernel vs the reference library on the same host + quick timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from convectionkernels_b200 import api, synth
from oracle.loader import Reference

R = Reference()
api.init(0)
print("rcp table", api.get_rcp_table()[:8], "threads", R.hardware_threads(), flush=True)
opt = api.Options()
plan = api.BC7EncodingPlan(); api.ConfigureBC7EncodingPlanFromQuality(plan, 100)
dplan = api.BC7EncodingPlan()

def modes(a): return np.bincount([(int(x[0]) & -int(x[0])).bit_length() - 1 for x in a], minlength=9)

ok = True
rb = synth.random_blocks_rgba8(4096, seed=3)
img = synth.image_to_blocks(synth.mixed_rgba8(1024, 1024))
for name, blocks, o, p in (("random q100", rb, opt, plan), ("random default-plan", rb, opt, dplan),
                           ("random better", rb, api.Options(flags=api.Flags.Better), plan),
                           ("random uniform", rb, api.Options(flags=api.Flags.Default | api.Flags.Uniform), plan),
                           ("mixed 1024^2 q100", img, opt, plan)):
    t = time.time(); ref = R.encode("BC7", blocks, np.frombuffer(bytes(memoryview(o)), np.uint8), np.frombuffer(p.tobytes(), np.uint8), threads=0); tr = time.time() - t
    t = time.time(); got = api.encode("BC7", blocks, o, p); tg = time.time() - t
    bad = int((ref != got).any(axis=1).sum())
    ok &= bad == 0
    print("%-22s blocks %7d mismatches %d  ref %.2fs (%.1f kblk/s)  gpu e2e %.3fs" % (name, len(blocks), bad, tr, len(blocks) / tr / 1e3, tg), modes(ref), flush=True)

# device-resident timing
for n in (65536, 262144, 1048576):
    src = synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))[:n]
    d = torch.from_numpy(src).cuda()
    out = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        api.encode("BC7", d, opt, plan, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        api.encode("BC7", d, opt, plan, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("n=%8d  %.2f ms  %.3f Mblocks/s" % (n, ms, n / ms / 1e3), flush=True)
print("PARITY", "OK" if ok else "FAILED")
