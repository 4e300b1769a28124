#!/usr/bin/env python3
"""Device-side timing of the decoders (CUDA events, inputs resident in HBM, L2 flushed between launches): blocks/s, algorithmic
GB/s (16 B read + 64 / 128 B written per block) and the fraction of the measured HBM copy bandwidth.
usage: time_decode.py [blocks]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from convectionkernels_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304          # 8192x8192 texels
api.init(0)
peak = 6457.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rng = np.random.default_rng(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for fmt, out_bytes in (("BC7", 64), ("BC6HU", 128), ("BC6HS", 128)):
    bc = rng.integers(0, 256, size=(n, 16), dtype=np.uint8)
    if fmt == "BC7":
        mode = rng.integers(0, 8, size=n)
        bc[:, 0] = (bc[:, 0] & ~(((1 << (mode + 1)) - 1) & 0xff) & 0xff) | (1 << mode)
    d = torch.from_numpy(bc).cuda()
    out = api.decode(fmt, d)
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        api.decode(fmt, d, out=out)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    gbs = n * (16 + out_bytes) / ms / 1e6
    print(json.dumps({"decode": fmt, "blocks": n, "ms": ms, "gblocks_per_s": n / ms / 1e6, "algorithmic_gb_per_s": gbs, "hbm_peak_gb_per_s": peak, "frac": gbs / peak}))
