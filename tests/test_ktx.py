"""The container step of the reference's sample packer (etc2packer/etc2packer.cpp:114-193, 277-284; etc2packer/ktxheader.h): the
68 bytes in front of the payload are checked on the CPU against a restatement of the sample's field assignments, the whole file
on the GPU against the unmodified reference encoding the sample's blocks."""
import struct

import numpy as np
import pytest

from convectionkernels_b200 import api, synth

# target -> (glInternalFormat, glBaseInternalFormat, bytes per block): ktxheader.h:9-33, etc2packer.cpp:147-180
TARGETS = {
    "ETC1": (0x8D64, 0x1907, 8),
    "ETC2": (0x9274, 0x1907, 8),
    "ETC2_RGBA": (0x9278, 0x1908, 16),
    "ETC2_PUNCHTHROUGH": (0x9276, 0x1908, 8),
    "EAC_R11U": (0x9270, 0x1903, 8),
    "EAC_R11S": (0x9271, 0x1903, 8),
}


def sample_header(fmt, w, h):
    internal, base, block_bytes = TARGETS[fmt]
    ident = bytes([0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A])
    # endianness, glType, glTypeSize, glFormat, internal, base, w, h, depth, array elements, faces (set to 0, then to 1: :139, :143),
    # mip levels, key-value bytes; then dataSize (:190-193)
    fields = [0x04030201, 0, 1, 0, internal, base, w, h, 0, 0, 1, 1, 0]
    return ident + struct.pack("<13I", *fields) + struct.pack("<I", ((w + 3) // 4) * ((h + 3) // 4) * block_bytes)


@pytest.mark.parametrize("fmt", sorted(TARGETS))
@pytest.mark.parametrize("w,h", [(4, 4), (70, 50), (4096, 4096), (1, 3)])
def test_header_bytes(fmt, w, h):
    got = api.ktx_header(fmt, w, h)
    assert len(got) == 68
    assert got == sample_header(fmt, w, h)


def test_header_rejects_what_the_sample_cannot_write():
    for fmt in ("BC7", "BC1", "ETC2_ALPHA"):
        with pytest.raises(api.CvttError) as e:
            api.ktx_header(fmt, 16, 16)
        assert e.value.status == -2
    with pytest.raises(api.CvttError):
        api.ktx_header("ETC2", 0, 16)


def _packer_blocks(img):
    h, w, _ = img.shape
    rows, groups = (h + 3) // 4, (w + 31) // 32
    ys = np.minimum(np.arange(rows * 4), h - 1)
    xs = np.minimum(np.arange(groups * 32), w - 1)
    return synth.image_to_blocks(img[ys][:, xs])


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["ETC2_RGBA", "ETC1", "ETC2_PUNCHTHROUGH", "EAC_R11U", "EAC_R11S"])
def test_packed_file_equals_the_sample_packer_run_on_the_reference(reference, fmt, tmp_path):
    import torch
    h, w = 50, 70
    img = np.ascontiguousarray(synth.mixed_rgba8(64, 96, seed=21)[:h, :w])
    o = api.Options()
    ob = np.frombuffer(bytes(memoryview(o)), np.uint8)
    blocks = _packer_blocks(img)
    if fmt.startswith("EAC_R11"):
        # etc2packer.cpp:232-237 (the signed target is fed normalizedUnsigned * 1023, as written there)
        total = blocks[..., :3].astype(np.float64).sum(axis=-1) / (255.0 * 3.0)
        blocks = np.floor(total * (2047.0 if fmt == "EAC_R11U" else 1023.0) + 0.5).astype(np.int16)
    want_all = reference.encode(fmt, np.ascontiguousarray(blocks), ob, threads=0)
    bb = TARGETS[fmt][2]
    rows, bpr, real = (h + 3) // 4, ((w + 31) // 32) * 8, (w + 3) // 4
    want = want_all.reshape(rows, bpr, bb)[:, :real].reshape(-1, bb)
    path = tmp_path / "out.ktx"
    api.write_ktx(str(path), fmt, torch.from_numpy(img).cuda(), o)
    data = path.read_bytes()
    assert data[:68] == sample_header(fmt, w, h)
    assert len(data) == 68 + want.size
    assert (np.frombuffer(data[68:], np.uint8).reshape(-1, bb) == want).all()
