#!/usr/bin/env python3
"""Workload for the ncu capture of the kernels outside the three search kernels: image tiler / untiler, decoders, BC1 / BC3 and
EAC alpha encoders.  Every call runs twice (the capture skips the first); sizes are larger than the 126 MB L2 where the kernel
is bandwidth-bound.  Run under: ncu --set full -k regex:'tile_image|untile_blocks|decode_kernel|s3tc_encode|eac_encode' ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from convectionkernels_b200 import api, synth

api.init(0)
o = api.Options()
img8 = torch.randint(0, 255, (8192, 8192, 4), device="cuda", dtype=torch.int32).to(torch.uint8)       # 256 MiB
img16 = torch.randint(0, 255, (4096, 8192, 4), device="cuda", dtype=torch.int32).to(torch.int16)      # 256 MiB
for img in (img8, img16):
    for _ in range(2):
        blocks = api.tile_image(img)
enc = torch.empty((api.tiled_block_count(8192, 8192), 16), dtype=torch.uint8, device="cuda")
for _ in range(2):
    api.untile_blocks(enc, 8192, 8192)
rgba = torch.from_numpy(synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))).cuda()               # 1 048 576 blocks, 64 MiB
rgba4 = torch.cat([rgba, rgba, rgba, rgba])                                                          # 4 194 304 blocks, 256 MiB
for fmt in ("BC1", "BC3", "ETC2_ALPHA"):
    for _ in range(2):
        api.encode(fmt, rgba4, o)
rng = np.random.default_rng(1)
bits = torch.from_numpy(rng.integers(0, 256, size=(1 << 22, 16), dtype=np.uint8)).cuda()           # 4 194 304 encoded blocks
for fmt in ("BC7", "BC6HU", "BC6HS"):
    for _ in range(2):
        api.decode(fmt, bits)
torch.cuda.synchronize()
