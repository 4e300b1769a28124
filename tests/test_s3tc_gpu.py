"""GPU parity tests for BC1-BC5 (run with -m gpu on the B200 box): the CUDA path through the C ABI against the golden vectors and
against the unmodified reference on the same host (oracle/_ref).  Bit-exact is the bar."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, first_mismatch
from convectionkernels_b200 import api, synth

pytestmark = pytest.mark.gpu
S3TC = ["BC1", "BC2", "BC3", "BC4U", "BC4S", "BC5U", "BC5S"]


def _opt_bytes(o):
    return np.frombuffer(bytes(memoryview(o)), np.uint8)


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)
    yield
    api.set_rcp_table(None)


@pytest.mark.parametrize("name", [n for p in ("bc1_", "bc2_", "bc3_", "bc4", "bc5") for n in golden_names(p)])
def test_golden(name):
    g = load_golden(name)
    api.set_rcp_table(g["rcp"])
    got = api.encode(str(g["fmt"]), g["blocks"], g["options"])
    api.set_rcp_table(None)
    assert (got == g["expected"]).all(), first_mismatch(g["expected"], got)


@pytest.mark.parametrize("fmt", S3TC)
def test_random_blocks_against_reference(reference, fmt):
    blocks = synth.random_blocks_rgba8(16384 + 8, seed=55)
    blocks[::2, :, 3] = np.random.default_rng(2).integers(0, 256, size=blocks[::2, :, 3].shape)
    for o in (api.Options(), None):
        if o is None:
            o = api.Options()
            o.flags = 0x208            # uniform weights, no paranoid metric
            o.refineRoundsS3TC = 3
            o.refineRoundsIIC = 2
            o.threshold = 0.25
        want = reference.encode(fmt, blocks, _opt_bytes(o), threads=0)
        got = api.encode(fmt, blocks, o)
        assert (got == want).all(), first_mismatch(want, got)


def test_config1_gradient_bc1(reference):
    """BASELINE.json configs[0]: EncodeBC1 on the 256x256 gradient"""
    blocks = synth.image_to_blocks(synth.gradient_rgba8(256, 256))
    o = api.Options()
    want = reference.encode("BC1", blocks, _opt_bytes(o), threads=0)
    got = api.EncodeBC1(blocks, o)
    assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("fmt", ["BC1", "BC2", "BC3"])
def test_exhaustive_against_reference(reference, fmt):
    """Flags::Better / Ultra: S3TC_Exhaustive (cluster-fit enumeration, single-colour tables, per-call maximum of the pixel counts)"""
    blocks = synth.random_blocks_rgba8(2048 + 8, seed=56)
    blocks[::2, :, 3] = np.random.default_rng(3).integers(0, 256, size=blocks[::2, :, 3].shape)
    for flags in (api.Flags.Better, api.Flags.Ultra, 0x288):
        o = api.Options()
        o.flags = flags
        want = reference.encode(fmt, blocks, _opt_bytes(o), threads=0)
        got = api.encode(fmt, blocks, o)
        assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("fmt", ["BC1", "BC3", "BC5S"])
def test_host_buffers_take_the_chunked_pipeline(fmt):
    """Host-buffer calls of the fast formats are cut into 131072-block chunks on two streams (cvtt_b200.cu); the result must not
    depend on it: two whole chunks plus a ragged tail of 37 groups against the one-launch device-pointer path."""
    import torch
    n = 131072 * 2 + 8 * 37
    blocks = synth.random_blocks_rgba8(n, seed=91)
    want = api.encode(fmt, torch.from_numpy(blocks).cuda(), api.Options()).cpu().numpy()
    got = api.encode(fmt, blocks, api.Options())
    assert (got == want).all(), first_mismatch(want, got)
    again = api.encode(fmt, blocks[: 131072 * 2 - 8], api.Options())          # below the threshold: single launch
    assert (again == want[: 131072 * 2 - 8]).all()
