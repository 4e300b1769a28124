import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from convectionkernels_b200 import api, synth
api.init(0)
blocks = synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))
hin = torch.from_numpy(blocks.reshape(-1)).pin_memory().numpy().reshape(blocks.shape)
o = api.Options()
for fmt in ["BC1", "BC3", "BC4U", "BC5U", "ETC2_ALPHA"]:
    want = api.encode(fmt, torch.from_numpy(blocks).cuda(), o).cpu().numpy()
    outb = want.shape[1]
    hout = torch.empty(want.shape, dtype=torch.uint8).pin_memory().numpy()
    api.encode(fmt, hin, o, out=hout)
    ok = bool((hout == want).all())
    t0 = time.perf_counter()
    for _ in range(5):
        api.encode(fmt, hin, o, out=hout)
    dt = (time.perf_counter() - t0) / 5
    # ragged size: not a multiple of the chunk
    n2 = 131072 * 2 + 8 * 37
    h2 = api.encode(fmt, hin[:n2], o)
    ok2 = bool((h2 == want[:n2]).all())
    print(fmt, "e2e pinned %.1f Mblocks/s" % (len(blocks) / dt / 1e6), "equal", ok, ok2)
