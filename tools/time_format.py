#!/usr/bin/env python3
"""Device-side timing of one format on the 4096x4096 synthetic texture of its BASELINE.json config (CUDA events, inputs resident in
HBM), plus the unmodified reference on a bounded sample for comparison.  usage: time_format.py FORMAT [blocks]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from convectionkernels_b200 import api, synth

fmt = sys.argv[1]
nblocks = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
api.init(0)
if fmt.startswith("BC6H"):
    blocks = synth.image_to_blocks(synth.hdr_ramp_f16(4096, 4096, signed=fmt.endswith("S")))[:nblocks]
else:
    blocks = synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))[:nblocks]
o, p = api.Options(), None
if fmt == "BC7":
    p = api.BC7EncodingPlan(); api.ConfigureBC7EncodingPlanFromQuality(p, 100)
d = torch.from_numpy(blocks).cuda()
out = api.encode(fmt, d, o, p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    api.encode(fmt, d, o, p, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
res = {"format": fmt, "blocks": int(blocks.shape[0]), "ms": ms, "mblocks_per_s": blocks.shape[0] / ms / 1e3}
try:
    from oracle.loader import Reference
    R = Reference()
    sample = blocks[:65536]
    ob = np.frombuffer(bytes(memoryview(o)), np.uint8)
    pb = None if p is None else np.frombuffer(p.tobytes(), np.uint8)
    t0 = time.perf_counter(); want = R.encode(fmt, sample, ob, pb, threads=0); dt = time.perf_counter() - t0
    res["reference_mblocks_per_s"] = len(sample) / dt / 1e6
    res["reference_threads"] = R.hardware_threads()
    res["bit_exact_on_sample"] = bool((out[:len(sample)].cpu().numpy() == want).all())
except Exception as e:
    res["reference"] = "unavailable: %s" % e
print(json.dumps(res))
