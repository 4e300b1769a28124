#!/usr/bin/env python3
"""Aggregate rate of T host threads that each make reference-sized BC7 calls (8 blocks, host buffers) at the same time -- what an
unmodified multi-threaded caller of the reference does.  usage: time_threads_small.py [blocks_per_call]"""
import os, sys, time, json, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from convectionkernels_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
api.init(0)
tex = synth.image_to_blocks(synth.mixed_rgba8(1024, 1024, seed=3))
o, p = api.Options(), api.BC7EncodingPlan()
api.ConfigureBC7EncodingPlanFromQuality(p, 100)
res = {"blocks_per_call": n, "split": os.environ.get("CVTTB200_BC7_SPLIT", "auto"), "kblocks_per_s_by_threads": {}}
CALLS = 200
for T in (1, 2, 4, 8, 16, 32):
    ins = [np.ascontiguousarray(tex[t * 512:t * 512 + n]) for t in range(T)]
    outs = [np.empty((n, 16), np.uint8) for _ in range(T)]
    for t in range(T):
        api.encode("BC7", ins[t], o, p, out=outs[t])

    def work(t):
        for _ in range(CALLS):
            api.encode("BC7", ins[t], o, p, out=outs[t])

    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t0
    res["kblocks_per_s_by_threads"][T] = round(T * CALLS * n / dt / 1e3, 1)
print(json.dumps(res))
