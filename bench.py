#!/usr/bin/env python3
"""Benchmark of the encode hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--format F] [--no-extras]

Headline (BASELINE.json configs[1]): EncodeBC7, ConfigureBC7EncodingPlanFromQuality(100), default cvtt::Options, on a synthetic
4096x4096 RGBA8 texture (1 048 576 4x4 blocks) per GPU.  A "step" is one pass of the hot path over the whole texture.  Rank 0
prints ONE JSON line:
  value         Mblocks/s, whole job (all ranks), inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e           the same metric through the public call with HOST (pinned) buffers: H2D copy + kernels + D2H copy per step
                (N > 1: plus the NCCL gather of the encoded ranges on rank 0 and the D2H copy of the gathered texture)
  roofline      the dominant kernel against the unit that binds it -- issue slots (warp instructions per second against
                SMs x 4 schedulers x the SM clock sampled during the run); the HBM view (algorithmic bytes / kernel time
                against the measured copy bandwidth) is carried as roofline.hbm: the path is compute-bound by four orders of
                magnitude, so that fraction is tiny by construction
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, all host threads) on the WHOLE texture, compared bit for bit with the
                GPU result (N = 1 only)
  other_configs BASELINE.json configs[2] (EncodeBC6HU, F16 HDR ramp) and configs[3] (EncodeETC2RGBA) measured the same way
                (N = 1 only, bounded steps)
  strong        BASELINE.json configs[4]: ONE 8192x8192 texture on rank 0, 8-block-group ranges scattered over NCCL, encoded,
                gathered back on rank 0 inside the step (strong scaling; reported at every N so that the curve has its base)
  latency_8block_ms   one cvtt::Kernels::EncodeBC7-sized call (8 blocks, host buffers) through the C ABI, next to
                reference_one_thread_8block_ms, the same call through the unmodified reference on one host thread
                (both also in every other_configs record);
                latency_ms_by_blocks_per_call gives the same for larger batches
With N > 1 (torchrun, one process per GPU) every rank encodes its own 4096x4096 texture (weak scaling) and the encoded
ranges are gathered on rank 0 with one NCCL gather inside the timed step.
`--impl reference` times the reference's own CPU implementation of the same configuration instead (rank 0 only).
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

SIDE = 4096
BLOCKS = (SIDE // 4) * (SIDE // 4)

# name -> (workload text, metric, input kind, bytes read per block, bytes written per block, dominant kernel, ncu summary)
FORMAT_CONFIGS = {
    "BC7": ("EncodeBC7 plan=FromQuality(100) Options=default 4096x4096 synthetic RGBA8 (1048576 blocks) per GPU", "Mblocks/s (4x4) BC7 q100", "rgba8", 64, 16, "bc7_encode_kernel<true,false>", "bc7"),
    "BC6HU": ("EncodeBC6HU Options=default (slow indexing) 4096x4096 synthetic F16 HDR ramp (1048576 blocks) per GPU", "Mblocks/s (4x4) BC6HU", "f16", 128, 16, "bc6h_encode_kernel<false,false>", "bc6hu"),
    "BC6HS": ("EncodeBC6HS Options=default 4096x4096 synthetic signed F16 HDR ramp per GPU", "Mblocks/s (4x4) BC6HS", "f16s", 128, 16, "bc6h_encode_kernel<true,false>", None),
    "ETC2_RGBA": ("EncodeETC2RGBA Options=default 4096x4096 synthetic RGBA8 (1048576 blocks) per GPU", "Mblocks/s (4x4) ETC2 RGBA", "rgba8", 64, 16, "etc_encode_kernel<2,false,false>", "etc2_rgba"),
    "ETC2": ("EncodeETC2 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) ETC2 RGB", "rgba8", 64, 8, "etc_encode_kernel<1,false,false>", None),
    "ETC1": ("EncodeETC1 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) ETC1", "rgba8", 64, 8, "etc_encode_kernel<0,false,false>", None),
    "ETC2_ALPHA": ("EncodeETC2Alpha 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) EAC alpha", "rgba8", 64, 8, "eac_encode_kernel<0>", None),
    "BC1": ("EncodeBC1 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC1", "rgba8", 64, 8, "s3tc_encode_kernel<BC1>", None),
    "BC2": ("EncodeBC2 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC2", "rgba8", 64, 16, "s3tc_encode_kernel<BC2>", None),
    "BC3": ("EncodeBC3 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC3", "rgba8", 64, 16, "s3tc_encode_kernel<BC3>", None),
    "BC4U": ("EncodeBC4U Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC4U", "rgba8", 64, 8, "s3tc_encode_kernel<BC4U>", None),
    "BC5U": ("EncodeBC5U Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC5U", "rgba8", 64, 16, "s3tc_encode_kernel<BC5U>", None),
}
OTHER_CONFIGS = ("BC6HU", "ETC2_RGBA")      # BASELINE.json configs[2], configs[3]
# sources whose content decides the instruction count of a format's dominant kernel (see kernel_source_hash)
KERNEL_SOURCES = {
    "bc7": ["bc7_core.cuh", "bc7_kernels.cu", "bc7_host.cpp", "cvtt_common.cuh"],
    "bc6hu": ["bc6h_core.cuh", "bc6h_kernels.cu", "cvtt_common.cuh"],
    "etc2_rgba": ["etc_core.cuh", "etc_kernels.cu", "cvtt_common.cuh"],
}

_TEXTURES = {}


def synthetic_blocks(kind, seed_offset=0):
    from convectionkernels_b200 import synth
    key = (kind, seed_offset)
    if key not in _TEXTURES:
        if kind == "f16":
            _TEXTURES[key] = synth.image_to_blocks(synth.hdr_ramp_f16(SIDE, SIDE, seed=99 + seed_offset))
        elif kind == "f16s":
            _TEXTURES[key] = synth.image_to_blocks(synth.hdr_ramp_f16(SIDE, SIDE, seed=99 + seed_offset, signed=True))
        else:
            _TEXTURES[key] = synth.image_to_blocks(synth.mixed_rgba8(SIDE, SIDE, seed=1234 + seed_offset))
    return _TEXTURES[key]


def make_config(fmt, world):
    """The workload description; identical for the GPU arm and the reference arm."""
    return {"workload": FORMAT_CONFIGS[fmt][0], "format": fmt, "blocks_per_gpu": BLOCKS,
            "parallelism": ("block-range shard x%d, NCCL gather of encoded ranges" % world) if world > 1 else "single GPU",
            "plan": "ConfigureBC7EncodingPlanFromQuality(100)" if fmt == "BC7" else None,
            "flags": "Default (BC7_FastIndexing|S3TC_Paranoid)",
            "l2": "GPU arm: 256 MiB flush write between timed iterations"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash(summary_name):
    """sha256 over the sources that make up a format's dominant kernel: the committed ncu summary records it at capture time,
    so a capture that no longer describes the code is recognised instead of silently reused."""
    h = hashlib.sha256()
    for name in KERNEL_SOURCES.get(summary_name, []):
        with open(os.path.join(ROOT, "convectionkernels_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def profile_constants(summary_name):
    """dram traffic per launch, instructions per block and issue-slot utilisation from the committed ncu capture (profiles/)."""
    if not summary_name:
        return {}
    try:
        with open(os.path.join(ROOT, "profiles", summary_name + "_kernel_ncu_summary.json")) as f:
            prof = json.load(f)
        prof["capture_matches_sources"] = prof.get("kernel_source_hash") == kernel_source_hash(summary_name)
        return prof
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (nvidia-smi's clocks line, via NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {}
            for n in dir(nv):
                if n.startswith("nvmlClocksThrottleReason") or n.startswith("nvmlClocksEventReason"):
                    v = getattr(nv, n)
                    if isinstance(v, int) and v not in (0,) and bin(v).count("1") == 1:
                        names.setdefault(v, n.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", ""))
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                except Exception:
                    r = 0
                for bit, name in names.items():
                    if r & bit and name not in ("GpuIdle", "None", "All"):
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:          # never let monitoring break the benchmark
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref = the unmodified sources compiled by
    oracle/Makefile), all host threads, a bounded sample of the configuration's texture per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.loader import Reference
    fmt = args.format
    workload, metric, kind = FORMAT_CONFIGS[fmt][:3]
    R = Reference()
    threads = R.hardware_threads()
    sample_blocks = 131072 if fmt in ("BC7", "BC6HU", "BC6HS") else 524288
    blocks = synthetic_blocks(kind)[:sample_blocks]
    opt, plan = R.default_options(), (R.plan_from_quality(100) if fmt == "BC7" else None)
    for _ in range(args.warmup):
        R.encode(fmt, blocks[:16384], opt, plan, threads=0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        R.encode(fmt, blocks, opt, plan, threads=0)
    dt = time.perf_counter() - t0
    v = sample_blocks * args.steps / dt / 1e6
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": v, "unit": "Mblocks/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(fmt, int(os.environ.get("WORLD_SIZE", str(args.gpus)))),
        "cpu_baseline": {"value": v, "unit": "Mblocks/s", "cores": threads, "kind": "reference",
                         "sample": "first %d blocks of the texture per step x %d steps, %d threads (a rate: the cost per block is content independent to a few per cent)" % (sample_blocks, args.steps, threads)},
        "e2e": {"value": v, "unit": "Mblocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


class Job:
    """Process-wide state of one bench run: rank, device, process group."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from convectionkernels_b200 import api
        self.torch, self.dist, self.api = torch, dist, api
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.distributed = self.world > 1
        if self.distributed:
            dist.init_process_group("nccl", device_id=self.dev)
        api.init(self.local_rank)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)       # > 126 MB L2
        self.sms = torch.cuda.get_device_properties(self.dev).multi_processor_count

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.distributed:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.distributed:
            self.dist.destroy_process_group()


def options_and_plan(api, fmt):
    opt, plan = api.Options(), None
    if fmt == "BC7":
        plan = api.BC7EncodingPlan()
        api.ConfigureBC7EncodingPlanFromQuality(plan, 100)
    return opt, plan


def measure_format(job, fmt, steps, warmup, with_cpu):
    """Device-timed leg, end-to-end leg, roofline and (N = 1) the whole-texture comparison with the reference for one format
    on this rank's 4096x4096 texture.  Returns the record on rank 0, None elsewhere."""
    torch, api = job.torch, job.api
    from convectionkernels_b200 import sharding
    workload, metric, kind, in_bytes, out_bytes, kernel_name, summary_name = FORMAT_CONFIGS[fmt]
    world, distributed, dev = job.world, job.distributed, job.dev
    warmup = max(warmup, 3)

    blocks_np = synthetic_blocks(kind, job.rank)
    host_in = torch.from_numpy(blocks_np.reshape(-1)).pin_memory()
    host_out = torch.empty(BLOCKS * out_bytes, dtype=torch.uint8).pin_memory()
    d_in = host_in.to(dev)
    d_out = torch.empty((BLOCKS, out_bytes), dtype=torch.uint8, device=dev)
    opt, plan = options_and_plan(api, fmt)
    total_blocks = BLOCKS * world

    def step():
        api.encode(fmt, d_in, opt, plan, out=d_out)
        if distributed:
            return sharding.gather_encoded(d_out, total_blocks, out_bytes, dst=0)
        return d_out

    for _ in range(warmup):
        step()
    job.barrier()

    # ---- device-timed leg -------------------------------------------------------------------------------
    sampler = ClockSampler(job.local_rank)
    sampler.start()
    kernel_events = []
    launches0 = api.launch_count()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ms_total = 0.0
    job.barrier()
    for _ in range(steps):
        job.flush.fill_(1)                               # L2 flush between timed iterations (outside the event pair)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record()
        k0.record()
        api.encode(fmt, d_in, opt, plan, out=d_out)
        k1.record()
        if distributed:
            sharding.gather_encoded(d_out, total_blocks, out_bytes, dst=0)
        e_end.record()
        torch.cuda.synchronize()
        step_ms_total += e_start.elapsed_time(e_end)
        kernel_events.append(k0.elapsed_time(k1))
    job.barrier()
    launches = api.launch_count() - launches0
    clocks = sampler.finish()

    total_ms = job.max_over_ranks(step_ms_total)
    value = total_blocks * steps / (total_ms / 1e3) / 1e6
    kernel_ms = float(np.mean(kernel_events))

    # ---- end-to-end leg: host buffers, copies inside the timed region ---------------------------------------
    host_out_np = host_out.numpy().reshape(BLOCKS, out_bytes)
    host_in_np = host_in.numpy()
    gathered_host = torch.empty(total_blocks * out_bytes, dtype=torch.uint8).pin_memory() if (distributed and job.rank == 0) else None

    def e2e_step():
        if not distributed:
            api.encode(fmt, host_in_np, opt, plan, out=host_out_np)       # the public call; returns when host_out is complete
            return
        d_in.copy_(host_in, non_blocking=True)
        api.encode(fmt, d_in, opt, plan, out=d_out)
        g = sharding.gather_encoded(d_out, total_blocks, out_bytes, dst=0)
        if job.rank == 0:
            gathered_host.copy_(g.reshape(-1), non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = job.max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_blocks * steps / e2e_s / 1e6
    if distributed:
        host_out.copy_(d_out.reshape(-1))
        same = True if job.rank != 0 else bool((gathered_host.numpy().reshape(total_blocks, out_bytes)[:BLOCKS] == host_out_np).all())
    else:
        same = bool((host_out_np == d_out.cpu().numpy()).all())

    if job.rank != 0:
        return None

    # ---- roofline ------------------------------------------------------------------------------------------
    peak, peak_src = measured_peak()
    algo_bytes = (in_bytes + out_bytes) * BLOCKS
    hbm_achieved = algo_bytes / (kernel_ms / 1e3) / 1e9
    prof = profile_constants(summary_name)
    hbm = {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "peak_source": peak_src,
           "algorithmic_bytes_per_launch": algo_bytes, "traffic": prof.get("dram_bytes_per_launch")}
    roofline = {"bound": "issue_slots", "achieved": None, "peak": None, "unit": "G warp-instructions/s", "frac": None,
                "traffic": prof.get("dram_bytes_per_launch"), "kernel": kernel_name, "kernel_ms": kernel_ms, "hbm": hbm,
                "note": "compute-bound path (millions of instructions per block against <= 144 bytes): the binding unit is the issue "
                        "slot; frac = instructions per block (committed ncu capture of this source) x blocks / 32 / kernel time / "
                        "(SMs x 4 schedulers x sampled SM clock)",
                "issue_slot_frac_ncu": prof.get("issue_slot_frac"), "fma_pipe_frac_ncu": prof.get("fma_pipe_frac")}
    ipb, clk = prof.get("warp_instructions_per_block"), clocks.get("sm_mhz")
    if ipb and clk and prof.get("capture_matches_sources"):
        roofline["achieved"] = ipb * BLOCKS / 32.0 / (kernel_ms / 1e3) / 1e9
        roofline["peak"] = job.sms * 4 * clk * 1e6 / 1e9
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
        roofline["instructions_per_block_ncu"] = ipb
    elif ipb:
        roofline["note"] += "; NOT computed: the committed capture was taken from different kernel sources (stale) or the clock could not be sampled"
        roofline["stale_capture"] = not prof.get("capture_matches_sources")

    # ---- CPU baseline: the unmodified reference on this box's host cores, the WHOLE texture, compared bit for bit ------
    cpu = None
    if with_cpu:
        try:
            from oracle.loader import Reference
            R = Reference()
            threads = R.hardware_threads()
            ob, pb = np.frombuffer(bytes(memoryview(opt)), np.uint8), (np.frombuffer(plan.tobytes(), np.uint8) if plan is not None else None)
            R.encode(fmt, blocks_np[:8192], ob, pb, threads=0)
            t0 = time.perf_counter()
            ref_out = R.encode(fmt, blocks_np, ob, pb, threads=0)
            dt = time.perf_counter() - t0
            differing = int((ref_out != host_out_np).any(axis=1).sum())
            cpu = {"value": BLOCKS / dt / 1e6, "unit": "Mblocks/s", "cores": threads, "kind": "reference",
                   "sample": "the whole texture (%d blocks), %d threads, %.1f s" % (BLOCKS, threads, dt),
                   "bit_exact_vs_gpu": differing == 0, "blocks_compared": BLOCKS, "blocks_differing": differing}
        except Exception as e:
            cpu = {"value": None, "unit": "Mblocks/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    else:
        cpu = {"value": None, "unit": "Mblocks/s", "cores": 0, "kind": "reference", "sample": "measured at N=1 only (the ranks share the host cores)"}

    return {"metric": metric, "value": value, "unit": "Mblocks/s", "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
            "kernel_ms": kernel_ms, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mblocks/s", "h2d_bytes_per_step": BLOCKS * in_bytes,
                    "d2h_bytes_per_step": (total_blocks if distributed else BLOCKS) * out_bytes,
                    "host_equals_device_result": same, "includes_gather": distributed},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}


def measure_strong(job, steps, warmup):
    """BASELINE.json configs[4]: EncodeBC7 quality 100 on ONE 8192x8192 RGBA8 texture (4 194 304 blocks) that lives on
    rank 0; per step the 8-block-group ranges are scattered to the ranks over NCCL, encoded, and the encoded ranges gathered
    on rank 0 (strong scaling: total work fixed).  Device-timed, max over ranks."""
    torch, api = job.torch, job.api
    from convectionkernels_b200 import synth, sharding
    dev, distributed = job.dev, job.distributed
    in_bytes, out_bytes = 64, 16
    side = 8192
    n_blocks = (side // 4) * (side // 4)
    d_all = None
    if job.rank == 0:
        def tile(r, c):                 # four different 4096x4096 textures; the first one is the headline texture
            if (r, c) == (0, 0):
                return synth.blocks_to_image(synthetic_blocks("rgba8", 0), SIDE, SIDE)
            return synth.mixed_rgba8(SIDE, SIDE, seed=1234 + 2 * r + c)
        img = np.concatenate([np.concatenate([tile(r, c) for c in range(2)], axis=1) for r in range(2)], axis=0)
        d_all = torch.from_numpy(synth.image_to_blocks(img).reshape(-1)).to(dev)
        del img
    opt, plan = options_and_plan(api, "BC7")

    def step():
        local = sharding.scatter_blocks(d_all, n_blocks, in_bytes, src=0, device=dev) if distributed else d_all
        enc = api.encode("BC7", local.reshape(-1, in_bytes), opt, plan)
        return sharding.gather_encoded(enc, n_blocks, out_bytes, dst=0) if distributed else enc

    out = None
    for _ in range(warmup):
        out = step()
    job.barrier()
    sampler = ClockSampler(job.local_rank)
    sampler.start()
    launches0 = api.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for _ in range(steps):
        job.flush.fill_(1)
        job.barrier()
        e0.record()
        out = step()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    job.barrier()
    launches = api.launch_count() - launches0
    clocks = sampler.finish()
    total_ms = job.max_over_ranks(total)
    if job.rank != 0:
        return None
    # the sharded result against ONE single-GPU encode of the whole texture
    single = api.encode("BC7", d_all.reshape(n_blocks, in_bytes), opt, plan)
    differing = int((single != out.reshape(n_blocks, out_bytes)).any(dim=1).sum().item())
    return {"metric": "Mblocks/s (4x4) BC7 q100", "value": n_blocks * steps / (total_ms / 1e3) / 1e6, "unit": "Mblocks/s", "scaling": "strong",
            "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "n_gpus": job.world, "blocks_total": n_blocks,
            "config": {"workload": "EncodeBC7 plan=FromQuality(100) Options=default ONE 8192x8192 synthetic RGBA8 (4194304 blocks) on rank 0, "
                                   "8-block-group ranges scattered / gathered over NCCL inside the step",
                       "parallelism": "block-range shard x%d" % job.world, "l2": "256 MiB flush write between timed iterations"},
            "clocks": clocks, "gpu_launches": int(launches),
            "sharded_equals_single_gpu": differing == 0, "blocks_compared": n_blocks}


def measure_latency(job, n_blocks=8, calls=30, fmt="BC7"):
    """One call of cvtt::Kernels::Encode* size (8 blocks) -- or a larger batch -- with host buffers through the C ABI
    (wall clock, ms per call)."""
    api = job.api
    opt, plan = options_and_plan(api, fmt)
    blocks = np.ascontiguousarray(synthetic_blocks(FORMAT_CONFIGS[fmt][2], job.rank)[4096:4096 + n_blocks])
    out = np.empty((n_blocks, FORMAT_CONFIGS[fmt][4]), np.uint8)
    for _ in range(3):
        api.encode(fmt, blocks, opt, plan, out=out)
    t0 = time.perf_counter()
    for _ in range(calls):
        api.encode(fmt, blocks, opt, plan, out=out)
    return (time.perf_counter() - t0) / calls * 1e3


def reference_latency(fmt, n_blocks=8, calls=5):
    """the same call through the unmodified reference on ONE host thread (oracle/_ref), ms per call; None without it"""
    try:
        from oracle.loader import Reference
        from convectionkernels_b200 import api
        R = Reference()
        opt, plan = options_and_plan(api, fmt)
        ob = np.frombuffer(bytes(memoryview(opt)), np.uint8)
        pb = None if plan is None else np.frombuffer(plan.tobytes(), np.uint8)
        blocks = np.ascontiguousarray(synthetic_blocks(FORMAT_CONFIGS[fmt][2], 0)[4096:4096 + n_blocks])
        R.encode(fmt, blocks, ob, pb, threads=1)
        t0 = time.perf_counter()
        for _ in range(calls):
            R.encode(fmt, blocks, ob, pb, threads=1)
        return (time.perf_counter() - t0) / calls * 1e3
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--format", default="BC7", choices=sorted(FORMAT_CONFIGS))
    ap.add_argument("--workload", default="texture4096", choices=["texture4096", "shard8192"],
                    help="texture4096: every rank encodes its own 4096x4096 texture (weak scaling, the default contract; the BC7 line "
                         "also carries the shard8192 record as `strong`); shard8192: only BASELINE.json configs[4]")
    ap.add_argument("--no-extras", action="store_true", help="headline only: no other_configs, strong and latency records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        reference_arm(args)
        return

    job = Job()
    if args.workload == "shard8192":
        rec = measure_strong(job, args.steps, max(args.warmup, 3))
        if job.rank == 0:
            rec.update({"higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic"})
            print(json.dumps(rec))
        job.close()
        return

    fmt = args.format
    main_rec = measure_format(job, fmt, args.steps, args.warmup, with_cpu=(job.world == 1))
    extras = fmt == "BC7" and not args.no_extras
    others, strong, latency = None, None, None
    if extras:
        latency = measure_latency(job)
        latency_curve = {str(n): measure_latency(job, n, calls=10) for n in (8, 64, 512, 4096, 32768, 262144)} if job.world == 1 else None
        if job.world == 1:
            others = {}
            for f in OTHER_CONFIGS:
                r = measure_format(job, f, max(2, min(args.steps, 3)), 3, with_cpu=True)
                r["config"] = make_config(f, 1)
                r["latency_8block_ms"] = measure_latency(job, fmt=f)
                r["reference_one_thread_8block_ms"] = reference_latency(f)
                others[f] = r
        strong = measure_strong(job, max(2, min(args.steps, 3)), 3)

    if job.rank == 0:
        line = {"metric": main_rec["metric"], "value": main_rec["value"], "unit": "Mblocks/s", "n_gpus": job.world, "steps": args.steps,
                "warmup": main_rec["warmup"], "ms_per_step": main_rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": make_config(fmt, job.world),
                "clocks": main_rec["clocks"], "e2e": main_rec["e2e"], "gpu_launches": main_rec["gpu_launches"],
                "roofline": main_rec["roofline"], "cpu_baseline": main_rec["cpu_baseline"]}
        if extras:
            line["latency_8block_ms"] = latency
            line["reference_one_thread_8block_ms"] = reference_latency("BC7") if job.world == 1 else None
            if latency_curve is not None:
                line["latency_ms_by_blocks_per_call"] = latency_curve      # host buffers, one call of that many blocks
            line["other_configs"] = others if others is not None else "measured at N=1 only"
            line["strong"] = strong
        print(json.dumps(line))
    job.close()


if __name__ == "__main__":
    main()
