"""GPU tests of the image <-> block-array kernels (the step before / after the encode path, reference
etc2packer/etc2packer.cpp:215-248 and :277-284): bit-exact against a numpy restatement of the sample packer's loops."""
import numpy as np
import pytest

from convectionkernels_b200 import api, synth

pytestmark = pytest.mark.gpu


def packer_blocks(img):
    """the sample packer's block extraction: groups of 8 blocks per 32 pixels, clamped coordinates"""
    h, w, c = img.shape
    rows, groups = (h + 3) // 4, (w + 31) // 32
    ys = np.minimum(np.arange(rows * 4), h - 1)
    xs = np.minimum(np.arange(groups * 32), w - 1)
    padded = img[ys][:, xs]
    return synth.image_to_blocks(padded)


@pytest.mark.parametrize("h,w", [(64, 64), (128, 96), (37, 53), (4, 4), (1, 1), (257, 1031)])
@pytest.mark.parametrize("dtype", [np.uint8, np.int16])
def test_tile_image_matches_the_sample_packer(h, w, dtype):
    import torch
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256 if dtype == np.uint8 else 30000, size=(h, w, 4)).astype(dtype)
    got = api.tile_image(torch.from_numpy(img).cuda()).cpu().numpy()
    want = packer_blocks(img)
    assert got.shape == want.shape == (api.tiled_block_count(w, h), 16, 4)
    assert (got == want).all()


def test_tile_encode_untile_equals_encoding_the_packed_blocks():
    import torch
    h, w = 100, 200          # neither a multiple of 4 / 32: padding blocks take part in the group semantics
    img = synth.mixed_rgba8(128, 256)[:h, :w]
    o, p = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(p, 30)
    d_blocks = api.tile_image(torch.from_numpy(np.ascontiguousarray(img)).cuda())
    enc = api.EncodeBC7(d_blocks, o, p)
    payload = api.untile_blocks(enc, w, h).cpu().numpy()
    want_all = api.EncodeBC7(packer_blocks(img), o, p)
    rows, bpr, real = (h + 3) // 4, ((w + 31) // 32) * 8, (w + 3) // 4
    want = want_all.reshape(rows, bpr, 16)[:, :real].reshape(-1, 16)
    assert (payload == want).all()
