#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libcvtt_ref.so).  Run in the dev container
(needs /root/reference to have been compiled by `make -C oracle ref`).  Each fixture stores the input PixelBlocks, the raw
cvtt::Options / cvtt::BC7EncodingPlan bytes, the reference's output blocks and the generating host's _mm_rcp_ps(0..16)
table (the reference's refinement step depends on that instruction; consumers replay it through set_rcp_table)."""
import os, sys, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.loader import Reference
from convectionkernels_b200 import synth

R = Reference()
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
rcp = R.rcp_table()


def options(flags=None, refine=None, weights=None, refine_bc6h=None, seeds=None, refine_s3tc=None, refine_iic=None, threshold=None):
    o = R.default_options().copy()
    if flags is not None:
        o[0:4] = np.frombuffer(struct.pack("<I", flags), np.uint8)
    if threshold is not None:
        o[4:8] = np.frombuffer(struct.pack("<f", threshold), np.uint8)
    if weights is not None:
        o[8:24] = np.frombuffer(struct.pack("<4f", *weights), np.uint8)
    if refine is not None:
        o[24:28] = np.frombuffer(struct.pack("<i", refine), np.uint8)
    for offset, value in ((28, refine_bc6h), (32, refine_iic), (36, refine_s3tc), (40, seeds)):
        if value is not None:
            o[offset:offset + 4] = np.frombuffer(struct.pack("<i", value), np.uint8)
    return o


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""


def save(name, fmt, blocks, opt, plan=None):
    if not name.startswith(ONLY):
        return
    out = R.encode(fmt, blocks, opt, plan)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), fmt=fmt, blocks=blocks, options=opt,
                        plan=plan if plan is not None else np.zeros(0, np.uint8), expected=out, rcp=rcp)
    print(name, fmt, blocks.shape, "->", out.shape)


rb = synth.random_blocks_rgba8(256, seed=11)
grad = synth.image_to_blocks(synth.gradient_rgba8(64, 64))          # config 1 content, 256 blocks
mixed = synth.image_to_blocks(synth.mixed_rgba8(128, 128))[:512]
q100, qdef, q40 = R.plan_from_quality(100), R.default_plan(), R.plan_from_quality(40)

save("bc7_random_q100", "BC7", rb, options(), q100)
save("bc7_random_defaultplan", "BC7", rb, options(), qdef)
save("bc7_random_q40", "BC7", rb, options(), q40)
save("bc7_random_better", "BC7", rb, options(flags=0x180), q100)                  # Flags::Better: no fast indexing
save("bc7_random_uniform", "BC7", rb, options(flags=0x208), q100)
save("bc7_random_refine3_weights", "BC7", rb, options(refine=3, weights=(1.0, 0.5, 2.0, 0.25)), q100)
save("bc7_random_refine1", "BC7", rb, options(refine=1), q100)
save("bc7_random_trysinglecolor", "BC7", rb, options(flags=0x118), q100)          # BC7_TrySingleColor | S3TC_Paranoid, exact indexing
save("bc7_gradient_q100", "BC7", grad, options(), q100)
save("bc7_mixed_q100", "BC7", mixed, options(), q100)
# config 1 (plumbing, CPU only): EncodeBC1 on the 256x256 gradient
save("bc1_gradient256", "BC1", synth.image_to_blocks(synth.gradient_rgba8(256, 256)), options())

# config 3 content: BC6H (unsigned / signed, slow and fast indexing, fewer rounds, uniform weights)
hu, hs = synth.random_blocks_f16(128, seed=31), synth.random_blocks_f16(128, seed=32, signed=True)
ramp = synth.image_to_blocks(synth.hdr_ramp_f16(64, 128))[:128]
save("bc6hu_random", "BC6HU", hu, options())
save("bc6hs_random", "BC6HS", hs, options())
save("bc6hu_random_fast", "BC6HU", hu, options(flags=0x148))
save("bc6hs_random_fast_uniform", "BC6HS", hs, options(flags=0x348))
save("bc6hu_random_rounds22", "BC6HU", hu, options(refine_bc6h=2, seeds=2))
save("bc6hs_random_rounds13_weights", "BC6HS", hs, options(refine_bc6h=1, seeds=3, weights=(1.0, 0.5, 2.0, 1.0)))
save("bc6hu_ramp", "BC6HU", ramp, options())

# config 4 content: ETC2 RGBA and the other ETC / EAC entry points (no reciprocal instruction on these paths)
save("etc2rgba_random", "ETC2_RGBA", rb, options())
save("etc2rgba_mixed", "ETC2_RGBA", mixed, options())
save("etc2_random_uniform", "ETC2", rb, options(flags=0x308))
save("etc2_gradient", "ETC2", grad, options())
save("etc1_random", "ETC1", rb, options())
save("etc1_mixed_uniform", "ETC1", mixed[:256], options(flags=0x308))
save("etc2alpha_mixed", "ETC2_ALPHA", mixed, options())
save("eacr11u_random", "EAC_R11U", synth.random_blocks_s16(256, seed=41), options())
save("eacr11s_random", "EAC_R11S", synth.random_blocks_s16(256, seed=42, signed=True), options())

# BC1-BC5 (config 1 content and the other S3TC entry points)
rba = rb.copy()
rba[::2, :, 3] = np.random.default_rng(5).integers(0, 256, size=rba[::2, :, 3].shape)
save("bc1_random_alpha", "BC1", rba, options())
save("bc1_random_noparanoid_refine3", "BC1", rba, options(flags=0x008, refine_s3tc=3))
save("bc2_random", "BC2", rba, options())
save("bc3_mixed", "BC3", mixed[:256], options())
save("bc3_random_uniform", "BC3", rba, options(flags=0x308))
save("bc4u_random", "BC4U", rba, options())
save("bc4s_random", "BC4S", rba, options())
save("bc5u_random_seeds2", "BC5U", rba, options(seeds=2, refine_iic=3))
save("bc5s_random", "BC5S", rba, options())
save("bc1_random_alpha_exhaustive", "BC1", rba[:128], options(flags=0x188))
save("bc3_random_exhaustive_uniform", "BC3", rba[:128], options(flags=0x288))
save("etc2rgba_random_bt709", "ETC2_RGBA", rb, options(flags=0x508))
save("etc1_random_bt709_accurate_uniform", "ETC1", rb, options(flags=0xF08))

# EncodeETC2PunchthroughAlpha: mixed alpha patterns, default / uniform / fake BT.709, thresholds inside and outside [0, 1]
pt = synth.punchthrough_blocks_rgba8(512, seed=41)
save("etc2punch_mixed", "ETC2_PUNCHTHROUGH", pt, options())
save("etc2punch_mixed_uniform_thr09", "ETC2_PUNCHTHROUGH", pt[:256], options(flags=0x308, threshold=0.9))
save("etc2punch_mixed_bt709_thr03", "ETC2_PUNCHTHROUGH", pt[:256], options(flags=0x508, threshold=0.3))
save("etc2punch_mixed_thr2", "ETC2_PUNCHTHROUGH", pt[:64], options(threshold=2.0))

# Flags::BC7_RespectPunchThrough (exact and fast indexing, with BC7_TrySingleColor): commits of modes 6 / 7 depend on the group
save("bc7_punch_respect_q100", "BC7", pt[:256], options(flags=0x128), q100)
save("bc7_punch_respect_fast_sc_q40", "BC7", pt[256:512], options(flags=0x1b8), q40)
