"""GPU parity tests for BC6H (run with -m gpu on the B200 box): the CUDA path through the C ABI against the golden vectors
and against the unmodified reference on the same host (oracle/_ref).  Bit-exact is the bar."""
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
    yield
    api.set_rcp_table(None)


@pytest.mark.parametrize("name", golden_names("bc6h"))
def test_golden(name):
    g = load_golden(name)
    api.set_rcp_table(g["rcp"])
    got = api.encode(str(g["fmt"]), g["blocks"], g["options"])
    api.set_rcp_table(None)
    assert (got == g["expected"]).all(), first_mismatch(g["expected"], got)


@pytest.mark.parametrize("fmt,flags", [("BC6HU", None), ("BC6HS", None), ("BC6HU", api.Flags.Default | 0x40), ("BC6HS", api.Flags.Default | 0x240)])
def test_random_blocks_against_reference(reference, fmt, flags):
    blocks = synth.random_blocks_f16(4096 + 8, seed=77, signed=fmt.endswith("S"))       # ragged last warp
    o = api.Options()
    if flags is not None:
        o.flags = flags
    want = reference.encode(fmt, blocks, _opt_bytes(o), threads=0)
    got = api.encode(fmt, blocks, o)
    assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("fmt,n", [("BC6HU", 8), ("BC6HU", 40), ("BC6HS", 1536), ("BC6HU", 37888), ("BC6HU", 37896), ("BC6HS", 37896)])
def test_small_calls_take_the_split_launch_and_match_the_reference(reference, fmt, n):
    """Calls of up to 37 888 blocks take the small-call launch (the 196 calls of the search dealt out to up to 49 times as many
    CTAs that record error histories, bc6h_resolve_kernel re-runs every block's winner call for its group); 37 896 blocks is
    the first size of the normal launch.  Bit-exact either way, group coupling included (random blocks: every group mixes)."""
    blocks = np.ascontiguousarray(np.concatenate([synth.image_to_blocks(synth.hdr_ramp_f16(256, 512, seed=12, signed=fmt.endswith("S"))),
                                                  synth.random_blocks_f16(37896 - 8192, seed=31, signed=fmt.endswith("S"))])[::-1][:n])
    o = api.Options()
    want = reference.encode(fmt, blocks, _opt_bytes(o), threads=0)
    got = api.encode(fmt, blocks, o)
    assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("fmt,flags", [("BC6HU", None), ("BC6HS", None), ("BC6HU", api.Flags.Default | 0x40), ("BC6HS", api.Flags.Default | 0x240)])
def test_every_slice_count_equals_the_normal_launch(fmt, flags):
    """75 776 random blocks in one call take the normal launch (verified against the reference above and in bench.py); the same
    blocks in calls of 8 ... 37 888 take the small-call launch with 196 ... 2 ranges.  Groups are independent, so every call
    must return the corresponding bytes of the big one -- random blocks, so every group exercises the mode-loop coupling."""
    blocks = synth.random_blocks_f16(75776, seed=91, signed=fmt.endswith("S"))
    o = api.Options()
    if flags is not None:
        o.flags = flags
    whole = api.encode(fmt, blocks, o)
    start = 0
    for n in (8, 16, 128, 640, 1536, 3072, 6144, 12288, 37888):
        got = api.encode(fmt, np.ascontiguousarray(blocks[start:start + n]), o)
        assert (got == whole[start:start + n]).all(), (n, first_mismatch(whole[start:start + n], got))
        start += n


def test_group_coupling_is_reproduced(reference):
    """the same block encodes differently next to different neighbours; the kernel must follow the reference"""
    base = synth.random_blocks_f16(64, seed=5)
    blocks = np.concatenate([np.concatenate([base[i:i + 1], base[8 * k:8 * k + 7]]) for k in range(8) for i in (3, 11)])
    o = api.Options()
    want = reference.encode("BC6HU", blocks, _opt_bytes(o))
    got = api.encode("BC6HU", blocks, o)
    assert (got == want).all(), first_mismatch(want, got)


def test_ramp_crop_and_device_pointers(reference):
    import torch
    blocks = synth.image_to_blocks(synth.hdr_ramp_f16(256, 512))           # 8192 blocks of the config-3 image
    o = api.Options()
    want = reference.encode("BC6HU", blocks, _opt_bytes(o), threads=0)
    host = api.EncodeBC6HU(blocks, o)
    dev = api.EncodeBC6HU(torch.from_numpy(blocks).cuda(), o)
    assert (host == want).all(), first_mismatch(want, host)
    assert (dev.cpu().numpy() == want).all()


def test_full_size_properties(reference):
    """BASELINE.json configs[2] size (4096x4096 F16): determinism, sub-range independence at group granularity, and a sample of
    groups against the reference."""
    import torch
    blocks = synth.image_to_blocks(synth.hdr_ramp_f16(4096, 4096))
    assert blocks.shape[0] == 1048576
    o = api.Options()
    d = torch.from_numpy(blocks).cuda()
    full = api.EncodeBC6HU(d, o).cpu().numpy()
    again = api.EncodeBC6HU(d, o).cpu().numpy()
    assert (full == again).all()
    for first, n in ((8 * 1001, 8 * 37), (524288, 4096)):
        part = api.EncodeBC6HU(blocks[first:first + n], o)
        assert (part == full[first:first + n]).all()
    rng = np.random.default_rng(3)
    groups = rng.choice(1048576 // 8, size=512, replace=False)
    idx = (groups[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
    want = reference.encode("BC6HU", blocks[idx], _opt_bytes(o), threads=0)
    assert (full[idx] == want).all(), first_mismatch(want, full[idx])
