"""GPU tests of cvttb200_decode (DecodeBC7 / DecodeBC6HU / DecodeBC6HS) through the C ABI against the unmodified reference, plus
encode -> decode round trips at full size."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from convectionkernels_b200 import api, synth
from test_decode_host import random_encoded_blocks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)


@pytest.mark.parametrize("fmt", ["BC7", "BC6HU", "BC6HS"])
@pytest.mark.parametrize("n", [8, 40, 65536 + 24])
def test_random_bit_patterns(reference, fmt, n):
    bc = random_encoded_blocks(fmt, n, seed=n)
    want = reference.decode(fmt, bc)
    got = api.decode(fmt, bc)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert (got == want).all()


def test_block_counts_that_are_not_multiples_of_the_warp():
    """decoding is per block: any count works and a prefix decodes to the prefix"""
    bc = random_encoded_blocks("BC7", 1000, seed=3)
    full = api.DecodeBC7(bc)
    for n in (1, 7, 33, 999):
        assert (api.DecodeBC7(bc[:n]) == full[:n]).all()
    assert api.DecodeBC7(bc[:0]).shape == (0, 16, 4)


@pytest.mark.parametrize("name", golden_names("bc7_") + golden_names("bc6h"))
def test_golden_encodings(reference, name):
    g = load_golden(name)
    fmt = str(g["fmt"])
    assert (api.decode(fmt, g["expected"]) == reference.decode(fmt, g["expected"])).all()


def test_device_pointers_and_errors():
    import torch
    bc = random_encoded_blocks("BC6HS", 4096, seed=5)
    d = torch.from_numpy(bc).cuda()
    got = api.DecodeBC6HS(d)
    assert got.is_cuda and got.dtype == torch.int16
    assert (got.cpu().numpy() == api.DecodeBC6HS(bc)).all()
    with pytest.raises(api.CvttError) as e:
        api.decode("ETC2", bc)
    assert e.value.status == -2          # the reference has no ETC / BC1-5 decoders


def test_bc7_round_trip_at_full_size(reference):
    """BASELINE.json configs[1] size: encode 4096x4096 on the GPU, decode on the GPU; the decode equals the reference's decode on a
    sample, and the round trip is close to the source (PSNR), which pins encoder and decoder against each other"""
    import torch
    blocks = synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))
    o, p = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, 40)
    d = torch.from_numpy(blocks).cuda()
    enc = api.EncodeBC7(d, o, p)
    dec = api.DecodeBC7(enc)
    sample = enc[:65536].cpu().numpy()
    assert (dec[:65536].cpu().numpy() == reference.decode("BC7", sample)).all()
    err = (dec.to(torch.float32) - d.to(torch.float32)) ** 2
    psnr = 10.0 * np.log10(255.0 ** 2 / float(err.mean()))
    assert psnr > 28.0, psnr          # the reference on the same content gives 30.5 dB (blue carries weight 0.1)


def test_bc6h_round_trip(reference):
    import torch
    blocks = synth.image_to_blocks(synth.hdr_ramp_f16(1024, 1024))
    o = api.Options()
    enc = api.EncodeBC6HU(torch.from_numpy(blocks).cuda(), o)
    dec = api.DecodeBC6HU(enc).cpu().numpy()
    assert (dec == reference.decode("BC6HU", enc.cpu().numpy())).all()
    src = blocks.view(np.float16).astype(np.float32)[..., :3]
    got = dec.view(np.float16).astype(np.float32)[..., :3]
    rel = np.abs(got - src) / np.maximum(np.abs(src), 1e-3)
    assert np.median(rel) < 0.02, float(np.median(rel))
