#!/usr/bin/env python3
"""Device-side time of BC7 calls of a few sizes (CUDA events, device buffers) next to the wall time of the same call with host
buffers.  CVTTB200_BC7_SPLIT=0 forces the normal launch, =n the small-call launch with n slices.  usage: time_small_calls.py [n ...]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from convectionkernels_b200 import api, synth

sizes = [int(a) for a in sys.argv[1:]] or [8, 32, 384, 1152, 2304, 4608, 9216, 18816]
api.init(0)
tex = synth.image_to_blocks(synth.mixed_rgba8(2048, 2048, seed=3))
o, p = api.Options(), api.BC7EncodingPlan()
api.ConfigureBC7EncodingPlanFromQuality(p, 100)
res = {"split": os.environ.get("CVTTB200_BC7_SPLIT", "auto"), "device_ms": {}, "host_call_ms": {}}
for n in sizes:
    b = np.ascontiguousarray(tex[:n])
    d = torch.from_numpy(b).cuda()
    out = api.encode("BC7", d, o, p)
    for _ in range(3):
        api.encode("BC7", d, o, p, out=out)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        api.encode("BC7", d, o, p, out=out)
    e1.record(); torch.cuda.synchronize()
    res["device_ms"][n] = round(e0.elapsed_time(e1) / reps, 4)
    ho = np.empty((n, 16), np.uint8)
    for _ in range(3):
        api.encode("BC7", b, o, p, out=ho)
    t0 = time.perf_counter()
    for _ in range(reps):
        api.encode("BC7", b, o, p, out=ho)
    res["host_call_ms"][n] = round((time.perf_counter() - t0) / reps * 1e3, 4)
print(json.dumps(res))
