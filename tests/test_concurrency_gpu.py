"""Concurrent callers (run with -m gpu): the reference has no global state and may be called from any number of threads
(/root/reference README.md:57).  libcvtt_b200 keeps per-call staging and scratch, so calls from different host threads overlap
on the device and stay bit-exact."""
import threading
import time

import numpy as np
import pytest

from convectionkernels_b200 import api, synth

pytestmark = pytest.mark.gpu

THREADS = 4
BLOCKS = 8192               # 256 warps = 22 CTAs of the BC7 kernel: four such calls fit the 148 SMs side by side
# The overlap tests ask for BC7_TrySingleColor: that search keeps the normal launch at every size.  Without it a call of this
# size takes the small-call launch, which fills all 148 SMs by itself (tests/test_bc7_gpu.py), so four of them at once can
# only queue on the device however well the library overlaps them.
def _overlap_options():
    return api.Options(flags=api.Options().flags | 0x10)
ROUNDS = 3


def _textures():
    return [np.ascontiguousarray(synth.image_to_blocks(synth.mixed_rgba8(512, 256, seed=50 + t))) for t in range(THREADS)]


def _plan(quality=100):
    plan = api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(plan, quality)
    return plan


def _run_threads(work):
    errors = []

    def guarded(t):
        try:
            work(t)
        except Exception as e:          # surfaced by the caller
            errors.append(e)

    threads = [threading.Thread(target=guarded, args=(t,)) for t in range(THREADS)]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    dt = time.perf_counter() - t0
    assert not errors, errors
    return dt


def test_host_buffer_calls_from_four_threads_overlap():
    """four host threads, each encoding its own texture through the plain host-buffer call (what an unmodified multi-threaded
    caller of the reference interface does): wall clock below 1.5x ONE thread's, results identical to the serial ones"""
    api.init(0)
    tex = _textures()
    assert tex[0].shape[0] == BLOCKS
    opt, plan = _overlap_options(), _plan()
    serial = [api.encode("BC7", tex[t], opt, plan) for t in range(THREADS)]      # also warms the plan cache and the pool
    outs = [np.empty((BLOCKS, 16), np.uint8) for _ in range(THREADS)]

    def one_thread_time():
        t0 = time.perf_counter()
        for _ in range(ROUNDS):
            api.encode("BC7", tex[0], opt, plan, out=outs[0])
        return time.perf_counter() - t0

    single = min(one_thread_time() for _ in range(3))

    def work(t):
        for _ in range(ROUNDS):
            api.encode("BC7", tex[t], opt, plan, out=outs[t])

    _run_threads(work)                                                            # warm-up: one stream per thread gets created
    wall = min(_run_threads(work) for _ in range(3))
    for t in range(THREADS):
        assert (outs[t] == serial[t]).all(), "thread %d differs from its serial result" % t
    assert wall < 1.5 * single, "4 threads x %d calls: %.1f ms, one thread: %.1f ms -- the calls serialised" % (ROUNDS, wall * 1e3, single * 1e3)


def test_device_pointer_calls_on_own_streams_overlap():
    """the same with device tensors, each thread on its own CUDA stream (enqueue-only calls)"""
    import torch
    api.init(0)
    tex = [torch.from_numpy(t).cuda() for t in _textures()]
    opt, plan = _overlap_options(), _plan()
    serial = [api.encode("BC7", tex[t], opt, plan).cpu().numpy() for t in range(THREADS)]
    outs = [torch.empty((BLOCKS, 16), dtype=torch.uint8, device="cuda") for _ in range(THREADS)]
    streams = [torch.cuda.Stream() for _ in range(THREADS)]
    torch.cuda.synchronize()

    def run(t):
        with torch.cuda.stream(streams[t]):
            for _ in range(ROUNDS):
                api.encode("BC7", tex[t], opt, plan, out=outs[t])
        streams[t].synchronize()

    def one_thread_time():
        t0 = time.perf_counter()
        run(0)
        return time.perf_counter() - t0

    single = min(one_thread_time() for _ in range(3))
    _run_threads(run)
    wall = min(_run_threads(run) for _ in range(3))
    for t in range(THREADS):
        assert (outs[t].cpu().numpy() == serial[t]).all()
    assert wall < 1.5 * single, "4 streams: %.1f ms, one stream: %.1f ms" % (wall * 1e3, single * 1e3)


def test_mixed_formats_and_plans_from_many_threads():
    """different formats, options and BC7 plans at the same time (plan cache look-ups, compiles and launches interleave);
    every result equals the one computed alone"""
    api.init(0)
    rgba = synth.random_blocks_rgba8(2048, seed=9)
    hdr = synth.image_to_blocks(synth.hdr_ramp_f16(128, 256, seed=4))
    jobs = []
    for q in (5, 35, 70, 100):
        jobs.append(("BC7", rgba, api.Options(), _plan(q)))
    o2 = api.Options()
    o2.flags = api.Flags.Default | 0x200
    jobs += [("ETC2_RGBA", rgba, api.Options(), None), ("ETC2", rgba, o2, None), ("BC6HU", hdr, api.Options(), None), ("BC3", rgba, api.Options(), None)]
    want = [api.encode(f, b, o, p) for f, b, o, p in jobs]
    got = [None] * len(jobs)

    def work(t):
        for k in range(t, len(jobs), THREADS):
            f, b, o, p = jobs[k]
            for _ in range(2):
                got[k] = api.encode(f, b, o, p)

    _run_threads(work)
    for k in range(len(jobs)):
        assert (got[k] == want[k]).all(), jobs[k][0]
