"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI (ctypes ->
cvttb200_encode), against the golden vectors, the plain-C oracle, and -- where oracle/_ref travelled -- the unmodified
reference on the same host.  Bit-exact is the bar everywhere."""
import ctypes
import os

import numpy as np
import pytest

from conftest import golden_names, load_golden, first_mismatch
from convectionkernels_b200 import api, synth

pytestmark = pytest.mark.gpu


def _opt_bytes(o):
    return np.frombuffer(bytes(memoryview(o)), np.uint8)


def _plan_bytes(p):
    return np.frombuffer(p.tobytes(), np.uint8)


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)
    yield
    api.set_rcp_table(None)


@pytest.mark.parametrize("name", golden_names("bc7_"))
def test_golden_vectors(name):
    g = load_golden(name)
    api.set_rcp_table(g["rcp"])
    try:
        got = api.encode("BC7", g["blocks"], g["options"], g["plan"])
    finally:
        api.set_rcp_table(None)
    assert (got == g["expected"]).all(), first_mismatch(g["expected"], got)


@pytest.mark.parametrize("flags,refine", [(api.Flags.Default, 2), (api.Flags.Better, 2), (api.Flags.Default | api.Flags.Uniform, 2),
                                          (api.Flags.Uniform, 1), (api.Flags.Default, 3), (api.Flags.Default, 0)])
def test_against_oracle_random(oracle, flags, refine):
    blocks = synth.random_blocks_rgba8(1024, seed=1000 + refine)
    o = api.Options(flags=flags, refineRoundsBC7=refine)
    for q in (100, 35):
        p = api.BC7EncodingPlan()
        api.ConfigureBC7EncodingPlanFromQuality(p, q)
        want = oracle.encode_bc7(blocks, _opt_bytes(o), _plan_bytes(p))
        got = api.encode("BC7", blocks, o, p)
        assert (got == want).all(), first_mismatch(want, got)


def test_against_oracle_default_constructed_plan(oracle):
    blocks = synth.random_blocks_rgba8(512, seed=31)
    o, p = api.Options(), api.BC7EncodingPlan()
    want = oracle.encode_bc7(blocks, _opt_bytes(o), _plan_bytes(p))
    got = api.encode("BC7", blocks, o, p)
    assert (got == want).all(), first_mismatch(want, got)


def test_edge_blocks_and_ragged_warps(oracle):
    """flat / extreme / transparent blocks; 8, 24 and 40 blocks (partial warps); groups whose votes differ inside a warp"""
    o, p = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, 100)
    special = np.zeros((40, 16, 4), np.uint8)
    special[1] = 255
    special[2, :, 3] = 255
    special[3, ::2] = 255
    special[4] = (12, 200, 77, 255)
    special[5] = (12, 200, 77, 128)
    special[6, :, :3] = np.arange(16)[:, None] * 17
    special[6, :, 3] = 255
    special[7, :, 3] = np.arange(16) * 17
    special[8:16] = synth.random_blocks_rgba8(8, seed=5)
    special[8:16, :, 3] = 0                      # whole group transparent: RGB modes disabled for it
    special[16:24] = synth.random_blocks_rgba8(8, seed=6)
    special[16:24, :, 3] = 255
    special[19, 7, 3] = 251                      # 250 < minAlpha < 255
    special[24:40] = synth.random_blocks_rgba8(16, seed=7)
    for n in (8, 24, 40):
        want = oracle.encode_bc7(special[:n], _opt_bytes(o), _plan_bytes(p))
        got = api.encode("BC7", special[:n], o, p)
        assert got.shape == (n, 16)
        assert (got == want).all(), first_mismatch(want, got)
    assert api.encode("BC7", special[:0], o, p).shape == (0, 16)


def test_argument_and_flag_errors():
    o, p = api.Options(), api.BC7EncodingPlan()
    with pytest.raises(api.CvttError) as e:
        api.encode("BC7", np.zeros((12, 16, 4), np.uint8), o, p)
    assert e.value.status == -1
    with pytest.raises(api.CvttError) as e:
        api.encode("BC7", np.zeros((8, 16, 4), np.uint8), o, None)
    assert e.value.status == -1                   # BC7 without a plan


@pytest.mark.parametrize("flags,quality", [(0x128, 100), (0x1a8, 40), (0x138, 70), (0x328, 100)])
def test_respect_punch_through_against_reference(reference, flags, quality):
    """Flags::BC7_RespectPunchThrough: the commits of modes 6 / 7 are masked per group with the reference's inverted AndNot
    (BC67.cpp:1406-1416), so results depend on the other blocks of the 8-block call; mixed opaque / binary / continuous alpha"""
    blocks = synth.punchthrough_blocks_rgba8(4096 + 24, seed=61)
    o, p = api.Options(), api.BC7EncodingPlan()
    o.flags = flags
    api.ConfigureBC7EncodingPlanFromQuality(p, quality)
    want = reference.encode("BC7", blocks, _opt_bytes(o), _plan_bytes(p), threads=0)
    got = api.EncodeBC7(blocks, o, p)
    assert (got == want).all(), first_mismatch(want, got)
    o.flags = flags & ~0x20
    assert (api.EncodeBC7(blocks, o, p) != got).any()          # the flag does change results on this input


def test_against_reference_on_this_host(reference):
    """1024x1024 crop of the headline image, unmodified reference with all host threads vs the GPU path"""
    blocks = synth.image_to_blocks(synth.mixed_rgba8(1024, 1024))
    o, p = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, 100)
    want = reference.encode("BC7", blocks, _opt_bytes(o), _plan_bytes(p), threads=0)
    got = api.encode("BC7", blocks, o, p)
    assert (got == want).all(), first_mismatch(want, got)
    # rcp table the library derived == the instruction the reference executes on this host
    assert (api.get_rcp_table()[1:] == reference.rcp_table()[1:]).all()


def test_device_pointers_match_host_pointers():
    import torch
    blocks = synth.random_blocks_rgba8(2048, seed=8)
    o, p = api.Options(), api.BC7EncodingPlan()
    host = api.encode("BC7", blocks, o, p)
    dev_in = torch.from_numpy(blocks).cuda()
    dev = api.encode("BC7", dev_in, o, p)
    assert dev.is_cuda
    assert (dev.cpu().numpy() == host).all()
    before = api.launch_count()
    api.encode("BC7", dev_in, o, p)
    assert api.launch_count() == before + 2


def test_full_size_properties(oracle):
    """BASELINE.json configs[1] size (4096x4096 -> 1 048 576 blocks): determinism, independence of 8-block groups
    (any aligned sub-range encodes to the same bytes), and a random sample of groups checked against the oracle."""
    import torch
    blocks = synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))
    assert blocks.shape[0] == 1048576
    o, p = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, 100)
    d = torch.from_numpy(blocks).cuda()
    full = api.encode("BC7", d, o, p).cpu().numpy()
    again = api.encode("BC7", d, o, p).cpu().numpy()
    assert (full == again).all()
    # mode census: the synthetic image must exercise every mode
    modes = np.bincount([(int(x) & -int(x)).bit_length() - 1 for x in full[:, 0]], minlength=8)
    assert (modes[:8] > 0).all(), modes
    # sub-range independence (what multi-GPU sharding relies on)
    for first, n in ((8 * 1001, 8 * 37), (524288, 4096), (1048576 - 64, 64)):
        part = api.encode("BC7", blocks[first:first + n], o, p)
        assert (part == full[first:first + n]).all()
    # sampled oracle comparison: 192 random groups
    rng = np.random.default_rng(12)
    groups = rng.choice(1048576 // 8, size=192, replace=False)
    idx = (groups[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
    want = oracle.encode_bc7(blocks[idx], _opt_bytes(o), _plan_bytes(p))
    assert (full[idx] == want).all(), first_mismatch(want, full[idx])


def test_two_lane_division_is_ieee():
    """f2_div (the packed reciprocal / Newton / remainder sequence of cvtt_common.cuh) against __fdiv_rn on 2^29 operand pairs"""
    assert api.selftest(samples=1 << 28, seed=7) == 0


@pytest.mark.parametrize("n", [8, 16, 40, 264, 1152, 1160, 2312, 4616, 9224, 18816, 18824, 28416, 28424])
def test_small_calls_take_the_split_launch_and_match_the_reference(reference, n):
    """Calls of up to 28 416 blocks take the small-call launch: the search of every block dealt out to 144 / 48 / 24 / 12 / 6 /
    3 / 2 CTAs (whichever still fits one wave) and the winners reduced by bc7_finish_kernel -- the reference's own call size,
    8 blocks, included; 28 424 is the first size of the normal launch.  Bit-exact either way."""
    blocks = synth.image_to_blocks(synth.mixed_rgba8(512, 1024, seed=77))[:n]
    o, p = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, 100)
    want = reference.encode("BC7", blocks, _opt_bytes(o), np.frombuffer(p.tobytes(), np.uint8), threads=0)
    got = api.EncodeBC7(blocks, o, p)
    assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("n,launches", [(65536, 4), (75576, 4), (81920, 2), (131072, 2)])
def test_partial_second_wave_is_sliced_and_matches_the_reference(reference, n, launches):
    """Calls of one to one and a third waves (148 SMs x 12 warps x 32 blocks; a 1024 x 1024 texture is 1.15): the CTAs behind
    the whole wave run as a sliced launch of their own (65 536 blocks: 23 CTAs x 6 slices).  Larger calls and calls whose second
    wave is more than a third full keep the single launch.  Bit-exact."""
    blocks = synth.image_to_blocks(synth.mixed_rgba8(1024, 2048, seed=79))[:n]
    o, p = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, 100)
    want = reference.encode("BC7", blocks, _opt_bytes(o), np.frombuffer(p.tobytes(), np.uint8), threads=0)
    before = api.launch_count()
    got = api.EncodeBC7(blocks, o, p)
    assert api.launch_count() == before + launches      # classification, whole wave (, sliced second wave, its finish kernel)
    assert (got == want).all(), first_mismatch(want, got)


@pytest.mark.parametrize("quality", [100, 55])
def test_every_slice_count_equals_the_normal_launch(quality):
    """65 536 random blocks in one call take the normal launch; the same blocks in calls of 8 ... 28 416 take the small-call launch
    with 144 ... 2 slices.  Groups are independent, so every call must return the corresponding bytes of the big one."""
    blocks = synth.random_blocks_rgba8(65536, seed=92)
    blocks[::3, :, 3] = 255                      # a third of the blocks opaque: groups of every class
    o, p = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, quality)
    whole = api.EncodeBC7(blocks, o, p)
    start = 0
    for n in (8, 384, 1152, 2304, 4608, 9216, 18816, 28416):
        got = api.EncodeBC7(np.ascontiguousarray(blocks[start:start + n]), o, p)
        assert (got == whole[start:start + n]).all(), (n, first_mismatch(whole[start:start + n], got))
        start += n


@pytest.mark.parametrize("flags,quality", [(0x208, 100), (0x008, 37), (0x000, 100), (0x000, 1)])
def test_small_calls_other_options(reference, flags, quality):
    """the small-call launch under Uniform | FastIndexing, FastIndexing with a sparse plan, slow indexing, the sparsest plan"""
    blocks = synth.image_to_blocks(synth.mixed_rgba8(128, 256, seed=78))[:1024]
    o, p = api.Options(), api.BC7EncodingPlan()
    o.flags = flags
    api.ConfigureBC7EncodingPlanFromQuality(p, quality)
    want = reference.encode("BC7", blocks, _opt_bytes(o), np.frombuffer(p.tobytes(), np.uint8), threads=0)
    got = api.EncodeBC7(blocks, o, p)
    assert (got == want).all(), first_mismatch(want, got)


def test_small_call_latency_beats_one_reference_thread(reference):
    """an unmodified caller's 8-block cvtt::Kernels::EncodeBC7 through the library must not be slower than the reference on one
    core (it was 2x slower before the small-call launch)"""
    import time
    blocks = synth.image_to_blocks(synth.mixed_rgba8(64, 64, seed=5))[:8]
    o, p = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, 100)
    ob, pb = _opt_bytes(o), np.frombuffer(p.tobytes(), np.uint8)
    out = np.empty((8, 16), np.uint8)
    for _ in range(3):
        api.encode("BC7", blocks, o, p, out=out)
    t0 = time.perf_counter()
    for _ in range(20):
        api.encode("BC7", blocks, o, p, out=out)
    ours = (time.perf_counter() - t0) / 20
    reference.encode("BC7", blocks, ob, pb, threads=1)
    t0 = time.perf_counter()
    for _ in range(5):
        reference.encode("BC7", blocks, ob, pb, threads=1)
    ref = (time.perf_counter() - t0) / 5
    assert ours < ref, "8-block call: %.2f ms here, %.2f ms for the reference on one core" % (ours * 1e3, ref * 1e3)
