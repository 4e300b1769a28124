"""cvttb200_encode_multi (one process, several GPUs, host buffers sharded by whole 8-block groups) against the single-device
call.  On a one-GPU box the device list degenerates to [0]; with two or more GPUs the ranges really land on different devices,
including a ragged split (group count not divisible by the device count) and a list that repeats the order backwards."""
import numpy as np
import pytest

from convectionkernels_b200 import api, synth

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("fmt", ["BC7", "BC6HU", "ETC2_RGBA", "BC1"])
def test_encode_multi_equals_single_device(fmt):
    api.init(0)
    n = 8 * 333                                  # 333 groups: never divisible by 2, 4 or 8 devices
    if fmt == "BC6HU":
        blocks = synth.image_to_blocks(synth.hdr_ramp_f16(256, 256, seed=5))[:n]
    else:
        blocks = synth.random_blocks_rgba8(n, seed=11)
    opt, plan = api.Options(), None
    if fmt == "BC7":
        plan = api.BC7EncodingPlan()
        api.ConfigureBC7EncodingPlanFromQuality(plan, 30)
    want = api.encode(fmt, blocks, opt, plan)
    count = _device_count()
    for devices in (None, [0], list(range(count))[::-1]):
        got = api.encode_multi(fmt, blocks, opt, plan, devices=devices)
        assert (got == want).all(), "devices=%r" % (devices,)


def test_encode_multi_argument_errors():
    api.init(0)
    blocks = synth.random_blocks_rgba8(16, seed=1)
    opt = api.Options()
    with pytest.raises(Exception):
        api.encode_multi("BC1", blocks, opt, devices=[0, 0])           # a device listed twice
    with pytest.raises(Exception):
        api.encode_multi("BC1", blocks, opt, devices=[_device_count()])  # out of range
    with pytest.raises(Exception):
        api.encode_multi("BC7", blocks, opt, None, devices=[0])        # BC7 needs a plan
    assert api.encode_multi("BC1", blocks[:0], opt).shape == (0, 8)


def test_encode_multi_overlaps_devices_with_pageable_buffers():
    """Ordinary (pageable) numpy buffers: a copy from or to pageable memory only returns when its data has moved, so the call
    enqueues every device's input and kernel before it collects any result.  With two devices the call must take clearly less
    than the one-device call."""
    import time
    count = _device_count()
    if count < 2:
        pytest.skip("needs two GPUs")
    for d in range(2):
        api.init(d)
    blocks = synth.image_to_blocks(synth.mixed_rgba8(2048, 2048, seed=8))          # 262144 blocks, ~35 ms of kernel on one B200
    opt, plan = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(plan, 100)
    out1, out2 = np.empty((len(blocks), 16), np.uint8), np.empty((len(blocks), 16), np.uint8)

    def timed(devices, out):
        api.encode_multi("BC7", blocks, opt, plan, devices=devices, out=out)       # warm-up: plan upload, pools
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            api.encode_multi("BC7", blocks, opt, plan, devices=devices, out=out)
            best = min(best, time.perf_counter() - t0)
        return best

    t1, t2 = timed([0], out1), timed([0, 1], out2)
    assert (out1 == out2).all()
    assert t2 < 0.75 * t1, "one device %.1f ms, two devices %.1f ms: the devices did not overlap" % (t1 * 1e3, t2 * 1e3)
