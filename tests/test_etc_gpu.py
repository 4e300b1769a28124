"""GPU parity tests for ETC1 / ETC2 / EAC (run with -m gpu on the B200 box): the CUDA path through the C ABI against the golden
vectors and against the unmodified reference on the same host (oracle/_ref).  Bit-exact is the bar."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, first_mismatch
from convectionkernels_b200 import api, synth

pytestmark = pytest.mark.gpu


def _opt_bytes(o):
    return np.frombuffer(bytes(memoryview(o)), np.uint8)


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)


@pytest.mark.parametrize("name", golden_names("etc") + golden_names("eac"))
def test_golden(name):
    g = load_golden(name)
    got = api.encode(str(g["fmt"]), g["blocks"], g["options"])
    assert (got == g["expected"]).all(), first_mismatch(g["expected"], got)


@pytest.mark.parametrize("fmt,flags", [("ETC2_RGBA", None), ("ETC2", api.Flags.Default | 0x200), ("ETC1", None), ("ETC1", api.Flags.Default | 0x200),
                                       ("ETC2_ALPHA", None)])
def test_random_blocks_against_reference(reference, fmt, flags):
    blocks = synth.random_blocks_rgba8(8192 + 24, seed=101)        # ragged last warp
    o = api.Options()
    if flags is not None:
        o.flags = flags
    want = reference.encode(fmt, blocks, _opt_bytes(o), threads=0)
    got = api.encode(fmt, blocks, o)
    assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("signed", [False, True])
def test_eac_r11_against_reference(reference, signed):
    blocks = synth.random_blocks_s16(65536, seed=7, signed=signed)
    o = api.Options()
    want = reference.encode("EAC_R11S" if signed else "EAC_R11U", blocks, _opt_bytes(o), threads=0)
    got = api.EncodeETC2Alpha11(blocks, signed, o)
    assert (got == want).all(), first_mismatch(want, got)


def test_weights_and_image_content_against_reference(reference):
    blocks = synth.image_to_blocks(synth.mixed_rgba8(512, 512))          # 16384 blocks of the config-4 image
    o = api.Options()
    o.redWeight, o.greenWeight, o.blueWeight = 1.0, 0.5, 0.25
    want = reference.encode("ETC2_RGBA", blocks, _opt_bytes(o), threads=0)
    got = api.EncodeETC2RGBA(blocks, o)
    assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("fmt", ["ETC2", "ETC2_RGBA", "ETC2_PUNCHTHROUGH"])
def test_alloc_time_options(reference, fmt):
    """cvtt::Kernels::AllocETC2Data fixes the chroma side axes from ITS options (ETC.cpp:3117-3145); the encode call's options
    give flags and error weights.  Different options at the two places, through the reference-name mirror."""
    blocks = synth.random_blocks_rgba8(2048, seed=41) if fmt != "ETC2_PUNCHTHROUGH" else synth.punchthrough_blocks_rgba8(2048, seed=5)
    enc, alloc = api.Options(), api.Options()
    enc.redWeight, enc.greenWeight, enc.blueWeight = 1.0, 0.5, 0.25
    alloc.redWeight, alloc.greenWeight, alloc.blueWeight = 0.1, 1.0, 0.7
    want = reference.encode(fmt, blocks, _opt_bytes(enc), threads=0, etc2_alloc_options=_opt_bytes(alloc))
    data = api.AllocETC2Data(alloc)
    call = {"ETC2": api.EncodeETC2, "ETC2_RGBA": api.EncodeETC2RGBA, "ETC2_PUNCHTHROUGH": api.EncodeETC2PunchthroughAlpha}[fmt]
    got = call(blocks, enc, data)
    api.ReleaseETC2Data(data)
    assert (got == want).all(), first_mismatch(want, got)
    if fmt == "ETC2":
        assert (want != reference.encode(fmt, blocks, _opt_bytes(enc), threads=0)).any()       # the test has teeth


def test_ultra_preset_is_accepted(reference):
    """Flags::Ultra carries ETC_FakeBT709Accurate without ETC_UseFakeBT709: the bit has no effect on its own"""
    blocks = synth.random_blocks_rgba8(1024, seed=3)
    o = api.Options()
    o.flags = api.Flags.Ultra
    want = reference.encode("ETC2_RGBA", blocks, _opt_bytes(o), threads=0)
    got = api.EncodeETC2RGBA(blocks, o)
    assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("fmt", ["ETC1", "ETC2", "ETC2_RGBA"])
@pytest.mark.parametrize("flags", [0x508, 0xD08, 0x708])
def test_fake_bt709_against_reference(reference, fmt, flags):
    """Flags::ETC_UseFakeBT709 (fast / accurate rounding, with Uniform)"""
    blocks = synth.random_blocks_rgba8(4096 + 8, seed=102)
    o = api.Options()
    o.flags = flags
    want = reference.encode(fmt, blocks, _opt_bytes(o), threads=0)
    got = api.encode(fmt, blocks, o)
    assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("flags,threshold", [(0x108, 0.5), (0x308, 0.9), (0x508, 0.3), (0xF08, 0.5), (0x108, 2.0), (0x108, -1.0)])
def test_punchthrough_against_reference(reference, flags, threshold):
    """EncodeETC2PunchthroughAlpha: mixed opaque / transparent / partly transparent blocks, whole groups of each, thresholds inside
    and outside [0, 1]; 8200 blocks = 16 CTAs of 512 plus a ragged tail, so CTAs mix groups that need different stages"""
    blocks = synth.punchthrough_blocks_rgba8(8200, seed=55)
    o = api.Options()
    o.flags, o.threshold = flags, threshold
    want = reference.encode("ETC2_PUNCHTHROUGH", blocks, _opt_bytes(o), threads=0)
    got = api.EncodeETC2PunchthroughAlpha(blocks, o)
    assert got.shape == (8200, 8)
    assert (got == want).all(), first_mismatch(want, got)


def test_punchthrough_opaque_image_and_stage_skipping(reference):
    """all-opaque input never enters the punch-through stages (the CTA-wide vote skips them) and equals what the reference produces;
    all-transparent input skips the opaque stages"""
    blocks = synth.random_blocks_rgba8(2048, seed=8).copy()
    blocks[:, :, 3] = 255
    o = api.Options()
    want = reference.encode("ETC2_PUNCHTHROUGH", blocks, _opt_bytes(o), threads=0)
    assert (api.EncodeETC2PunchthroughAlpha(blocks, o) == want).all()
    blocks[:, :, 3] = 0
    want = reference.encode("ETC2_PUNCHTHROUGH", blocks, _opt_bytes(o), threads=0)
    assert (api.EncodeETC2PunchthroughAlpha(blocks, o) == want).all()


def test_full_size_properties(reference):
    """BASELINE.json configs[3] size (4096x4096 RGBA8 -> ETC2 RGBA): determinism, sub-range independence at group granularity, and a
    sample of groups against the reference."""
    import torch
    blocks = synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))
    o = api.Options()
    d = torch.from_numpy(blocks).cuda()
    full = api.EncodeETC2RGBA(d, o).cpu().numpy()
    again = api.EncodeETC2RGBA(d, o).cpu().numpy()
    assert (full == again).all()
    for first, n in ((8 * 1001, 8 * 37), (524288, 4096)):
        part = api.EncodeETC2RGBA(blocks[first:first + n], o)
        assert (part == full[first:first + n]).all()
    rng = np.random.default_rng(4)
    groups = rng.choice(1048576 // 8, size=2048, replace=False)
    idx = (groups[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
    want = reference.encode("ETC2_RGBA", blocks[idx], _opt_bytes(o), threads=0)
    assert (full[idx] == want).all(), first_mismatch(want, full[idx])
