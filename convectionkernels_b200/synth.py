"""Seeded synthetic inputs for the five BASELINE.json configurations.

Images are converted to the reference's PixelBlock arrays in the etc2packer order
(reference etc2packer/etc2packer.cpp:215-248): row-major 4x4 blocks, 8 horizontally consecutive blocks = one
reference call ("group"), pixels row-major inside a block, RGBA per pixel.
"""
import numpy as np


def image_to_blocks(img):
    """(H, W, C) image -> (H/4 * W/4, 16, C) PixelBlock array (H, W multiples of 4)."""
    h, w, c = img.shape
    assert h % 4 == 0 and w % 4 == 0
    b = img.reshape(h // 4, 4, w // 4, 4, c).transpose(0, 2, 1, 3, 4)
    return np.ascontiguousarray(b.reshape((h // 4) * (w // 4), 16, c))


def blocks_to_image(blocks, h, w):
    c = blocks.shape[-1]
    b = blocks.reshape(h // 4, w // 4, 4, 4, c).transpose(0, 2, 1, 3, 4)
    return np.ascontiguousarray(b.reshape(h, w, c))


def gradient_rgba8(h=256, w=256):
    """Config 1: R=x, G=y, B=(x+y)/2, A=255."""
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([x & 255, y & 255, ((x + y) // 2) & 255, np.full_like(x, 255)], axis=-1).astype(np.uint8)
    return img


def mixed_rgba8(h=4096, w=4096, seed=1234):
    """Configs 2/4/5: smooth gradients + seeded noise + hard edges; alpha = 255 on 7/8 of the 8-block groups, an
    alpha ramp on every 8th group, plus scattered single translucent / punch-through blocks, so that every BC7 mode
    and both cross-lane group flags (reference ConvectionKernels_BC67.cpp:1069,1072) are exercised."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.int32)
    base = np.stack([
        (x * 255) // max(w - 1, 1),
        (y * 255) // max(h - 1, 1),
        (((x // 4) ^ (y // 4)) * 5) & 255,
    ], axis=-1).astype(np.int32)
    # low-frequency colour variation
    base[..., 0] = (base[..., 0] + 40 * np.sin(y / 37.0) + 25 * np.cos(x / 53.0)).astype(np.int32)
    base[..., 2] = (base[..., 2] // 2 + 60 * np.sin((x + y) / 91.0) + 64).astype(np.int32)
    # noise whose amplitude varies by 64x64 region (0 .. 48)
    amp = rng.integers(0, 49, size=(h // 64 + 1, w // 64 + 1))
    amp = np.repeat(np.repeat(amp, 64, axis=0), 64, axis=1)[:h, :w]
    noise = rng.integers(-128, 129, size=(h, w, 3))
    img = base + (noise * amp[..., None]) // 128
    # hard edges: diagonal stripes and rectangles in some regions
    stripe = ((x + 2 * y) // 23) % 7 == 0
    img[stripe] = 255 - img[stripe]
    rect = ((x // 96) % 3 == 1) & ((y // 80) % 4 == 2) & (((x % 96) < 40) ^ ((y % 80) < 30))
    img[rect] = img[rect] // 4
    img = np.clip(img, 0, 255)

    alpha = np.full((h, w), 255, dtype=np.int32)
    bx, by = x // 4, y // 4
    group = by * (w // 4 // 8 if w >= 32 else 1) + bx // 8
    ramp_group = (group % 8) == 7
    ramp = ((x * 3 + y * 5) & 255)
    alpha[ramp_group] = ramp[ramp_group]
    # scattered single blocks: translucent noise, punch-through, nearly opaque (251..255)
    nb = (h // 4) * (w // 4)
    kind = rng.integers(0, 400, size=nb).reshape(h // 4, w // 4)
    kind_px = np.repeat(np.repeat(kind, 4, axis=0), 4, axis=1)
    a_noise = rng.integers(0, 256, size=(h, w))
    m = kind_px == 0
    alpha[m] = a_noise[m]
    m = kind_px == 1
    alpha[m] = np.where(a_noise[m] < 128, 0, 255)
    m = kind_px == 2
    alpha[m] = 251 + (a_noise[m] % 5)
    out = np.concatenate([img, alpha[..., None]], axis=-1).astype(np.uint8)
    return out


def random_blocks_rgba8(n, seed=0, kind="mix"):
    """Small adversarial block sets for parity tests: pure noise, flat, two-colour, gradients, alpha variants."""
    rng = np.random.default_rng(seed)
    blocks = np.zeros((n, 16, 4), dtype=np.uint8)
    for i in range(n):
        k = i % 8 if kind == "mix" else kind
        if k == 0:    # noise
            b = rng.integers(0, 256, size=(16, 4))
        elif k == 1:  # flat colour + tiny noise
            b = rng.integers(0, 256, size=(1, 4)) + rng.integers(-2, 3, size=(16, 4))
        elif k == 2:  # two clusters
            c = rng.integers(0, 256, size=(2, 4))
            b = c[rng.integers(0, 2, size=16)] + rng.integers(-6, 7, size=(16, 4))
        elif k == 3:  # linear gradient
            c0, c1 = rng.integers(0, 256, size=(2, 4))
            t = (np.arange(16) % 4 + np.arange(16) // 4) / 6.0
            b = c0[None, :] + (c1 - c0)[None, :] * t[:, None] + rng.integers(-3, 4, size=(16, 4))
        elif k == 4:  # exactly flat
            b = np.repeat(rng.integers(0, 256, size=(1, 4)), 16, axis=0)
        elif k == 5:  # three clusters
            c = rng.integers(0, 256, size=(3, 4))
            b = c[rng.integers(0, 3, size=16)] + rng.integers(-4, 5, size=(16, 4))
        elif k == 6:  # extremes
            b = rng.choice([0, 255], size=(16, 4))
        else:         # smooth noise
            b = rng.integers(96, 160, size=(16, 4))
        b = np.clip(b, 0, 255)
        amode = rng.integers(0, 6)
        if amode <= 2:
            b[:, 3] = 255
        elif amode == 3:
            b[:, 3] = rng.choice([0, 255], size=16)
        elif amode == 4:
            b[:, 3] = 251 + rng.integers(0, 5, size=16)
        blocks[i] = b
    return blocks


def punchthrough_blocks_rgba8(n, seed=0):
    """RGBA8 blocks for EncodeETC2PunchthroughAlpha: the colour content of random_blocks_rgba8 with alpha patterns that exercise
    every branch of the reference's punch-through flow (ETC.cpp:1664-1887): opaque blocks, fully transparent blocks, binary and
    continuous alpha, a single hole, a transparent half, plus whole 8-block groups that are all opaque / all transparent."""
    rng = np.random.default_rng(seed)
    b = random_blocks_rgba8(n, seed=seed).copy()
    kind = rng.integers(0, 6, size=n)
    group = np.arange(n) // 8
    kind[group % 7 == 0] = 0
    kind[group % 7 == 1] = 1
    for i in range(n):
        k = kind[i]
        if k == 0:
            b[i, :, 3] = 255
        elif k == 1:
            b[i, :, 3] = rng.integers(0, 100)
        elif k == 2:
            b[i, :, 3] = rng.integers(0, 2, size=16) * 255
        elif k == 3:
            b[i, :, 3] = rng.integers(0, 256, size=16)
        elif k == 4:
            b[i, :, 3] = 255
            b[i, rng.integers(0, 16), 3] = 0
        else:
            b[i, :8, 3] = 0
            b[i, 8:, 3] = 255
    return b


def hdr_ramp_f16(h=4096, w=4096, seed=99, signed=False):
    """Config 3: HDR ramp R=0.01*2^(x/64), G=0.02*2^(y/64), B=0.5+noise; returns int16 half bit patterns, RGBA."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    r = 0.01 * np.exp2((x % 1024) / 64.0)
    g = 0.02 * np.exp2((y % 1024) / 64.0)
    b = 0.5 + 0.25 * np.sin(x / 17.0) * np.cos(y / 29.0) + rng.random((h, w), dtype=np.float32) * 0.1
    img = np.stack([r, g, b, np.ones_like(r)], axis=-1)
    if signed:
        neg = rng.random((h, w)) < 0.05
        img[neg, :3] *= -1.0
    img = np.clip(img, -65000.0, 65000.0).astype(np.float16)
    return img.view(np.int16)


def random_blocks_f16(n, seed=0, signed=False):
    """n PixelBlockF16 blocks (int16 half bit patterns, RGBA) covering flat, noisy, ramp, two-colour, wide-range and
    (signed) negative content, so that BC6H single- and two-subset modes of every precision get chosen."""
    rng = np.random.default_rng(seed)
    b = np.zeros((n, 16, 4), np.float32)
    for i in range(n):
        kind = i % 6
        base = rng.uniform(0, 4, 3) * rng.choice([0.01, 0.1, 1, 10, 100])
        if kind == 0:
            px = base[None, :] * np.ones((16, 1))
        elif kind == 1:
            px = base[None, :] * (1 + rng.uniform(-0.3, 0.3, (16, 3)))
        elif kind == 2:
            g = np.linspace(0, 1, 16)[:, None]
            px = base[None, :] * (0.2 + g) + rng.uniform(0, 0.02, (16, 3))
        elif kind == 3:
            m = (rng.random(16) < 0.5)[:, None]
            px = np.where(m, base[None, :], rng.uniform(0, 2, 3)[None, :]) * (1 + rng.uniform(-0.05, 0.05, (16, 3)))
        elif kind == 4:
            px = rng.uniform(0, 65000, (16, 3)) * rng.choice([1, 1e-3, 1e-6])
        else:
            px = rng.normal(0, 1, (16, 3)) * base[None, :]
        if not signed and kind != 5:
            px = np.abs(px)
        b[i, :, :3] = px
        b[i, :, 3] = 1
    return np.ascontiguousarray(b.astype(np.float16).view(np.int16))


def random_blocks_s16(n, seed=0, signed=False):
    """n PixelBlockScalarS16 blocks (int16 [16]) for EncodeETC2Alpha11: full-range noise, quantised steps, narrow ranges and
    out-of-range values (the encoder clamps to 0..2047 / -1023..1023)."""
    rng = np.random.default_rng(seed)
    b = rng.integers(-1100 if signed else -50, 2100, size=(n, 16)).astype(np.int16)
    b[::3] = (b[::3] // 64) * 8
    k = len(b[1::5])
    b[1::5] = b[1::5, :1] + rng.integers(-20, 20, size=(k, 16))
    return np.ascontiguousarray(b)
