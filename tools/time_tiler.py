#!/usr/bin/env python3
"""Device-side timing of the image -> block-array kernel on 8192x8192 RGBA8 and 4096x4096 RGBA16F (inputs larger than the
126 MB L2), CUDA events, against the measured copy bandwidth of MEASURED_PEAKS.json.  Prints one JSON line per case."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from convectionkernels_b200 import api

api.init(0)
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
for name, h, w, dtype in (("rgba8_8192", 8192, 8192, torch.uint8), ("rgba16f_8192x4096", 4096, 8192, torch.int16)):
    img = torch.randint(0, 255, (h, w, 4), device="cuda", dtype=torch.int32).to(dtype)
    out = api.tile_image(img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        api.tile_image(img, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = 2 * img.numel() * img.element_size()          # read the image once, write the blocks once
    # the same bytes through torch's copy kernel, for reference
    dst = torch.empty_like(img)
    e0.record()
    for _ in range(reps):
        dst.copy_(img)
    e1.record(); torch.cuda.synchronize()
    ms_copy = e0.elapsed_time(e1) / reps
    print(json.dumps({"case": name, "ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6, "peak_gbs": peak,
                      "frac": nbytes / ms / 1e6 / peak, "torch_copy_gbs": nbytes / ms_copy / 1e6}))
