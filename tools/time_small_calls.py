#!/usr/bin/env python3
"""Device-side time of BC7 calls of a few sizes (CUDA events, device buffers) next to the wall time of the same call with host
buffers.  CVTTB200_BC7_SPLIT=0 forces the normal launch, =n the small-call launch with n slices.  usage: time_small_calls.py [n ...]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from convectionkernels_b200 import api, synth

fmt = "BC7"
args = sys.argv[1:]
if args and not args[0].isdigit():
    fmt = args.pop(0)
sizes = [int(a) for a in args] or [8, 32, 384, 1152, 2304, 4608, 9216, 18816]
api.init(0)
if fmt.startswith("BC6H"):
    tex = synth.image_to_blocks(synth.hdr_ramp_f16(2048, 2048, signed=fmt.endswith("S")))
else:
    tex = synth.image_to_blocks(synth.mixed_rgba8(2048, 2048, seed=3))
o, p = api.Options(), None
if fmt == "BC7":
    p = api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, 100)
res = {"format": fmt, "split": os.environ.get("CVTTB200_BC7_SPLIT", "auto"), "device_ms": {}, "host_call_ms": {}}
try:
    from oracle.loader import Reference
    R = Reference()
    ob = np.frombuffer(bytes(memoryview(o)), np.uint8)
    pb = None if p is None else np.frombuffer(p.tobytes(), np.uint8)
    sample = np.ascontiguousarray(tex[:8])
    R.encode(fmt, sample, ob, pb, threads=1)
    t0 = time.perf_counter()
    for _ in range(5):
        R.encode(fmt, sample, ob, pb, threads=1)
    res["reference_one_thread_8_blocks_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 4)
except Exception as e:
    res["reference"] = "unavailable: %s" % e
for n in sizes:
    b = np.ascontiguousarray(tex[:n])
    d = torch.from_numpy(b).cuda()
    out = api.encode(fmt, d, o, p)
    for _ in range(3):
        api.encode(fmt, d, o, p, out=out)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        api.encode(fmt, d, o, p, out=out)
    e1.record(); torch.cuda.synchronize()
    res["device_ms"][n] = round(e0.elapsed_time(e1) / reps, 4)
    ho = np.empty_like(out.cpu().numpy())
    for _ in range(3):
        api.encode(fmt, b, o, p, out=ho)
    t0 = time.perf_counter()
    for _ in range(reps):
        api.encode(fmt, b, o, p, out=ho)
    res["host_call_ms"][n] = round((time.perf_counter() - t0) / reps * 1e3, 4)
print(json.dumps(res))
