#!/usr/bin/env python3
"""Benchmark of the encode hot path: BC7, ConfigureBC7EncodingPlanFromQuality(100), default cvtt::Options, on a
synthetic 4096x4096 RGBA8 texture (1 048 576 4x4 blocks) per GPU -- BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over the whole texture.  Prints ONE JSON line (rank 0).
  value      Mblocks/s, whole job (all ranks), inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e        same metric through the public call with HOST (pinned) buffers: H2D copy + kernel + D2H copy per step
  roofline   HBM view of the kernel: algorithmic bytes (64 B read + 16 B written per block) / kernel time vs the
             measured copy bandwidth.  The path is compute-bound by ~4 orders of magnitude, so `frac` is
             tiny by construction; `issue_slot_frac_ncu` (from the committed ncu capture) is the fraction that matters.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, all host threads) on a bounded sample of the same texture
With N > 1 (torchrun, one process per GPU) every rank encodes its own 4096x4096 texture (weak scaling) and the encoded
ranges are gathered on rank 0 with one NCCL gather inside the timed step.
`--impl reference` times the reference's own CPU implementation instead (rank 0 only).
"""
import argparse
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
CPU_SAMPLE_BLOCKS = 262144               # first 1024 rows of the texture

# The headline (default) is BASELINE.json configs[1], BC7.  --format selects one of the other single-GPU configurations
# (configs[2] BC6HU, configs[3] ETC2_RGBA) or any other implemented format; the JSON contract is the same.
# name -> (workload text, metric, input kind, bytes read per block, bytes written per block, dominant kernel)
FORMAT_CONFIGS = {
    "BC7": ("EncodeBC7 plan=FromQuality(100) Options=default 4096x4096 synthetic RGBA8 (1048576 blocks) per GPU", "Mblocks/s (4x4) BC7 q100", "rgba8", 64, 16, "bc7_encode_kernel<true,false>"),
    "BC6HU": ("EncodeBC6HU Options=default (slow indexing) 4096x4096 synthetic F16 HDR ramp (1048576 blocks) per GPU", "Mblocks/s (4x4) BC6HU", "f16", 128, 16, "bc6h_encode_kernel<false,false>"),
    "BC6HS": ("EncodeBC6HS Options=default 4096x4096 synthetic signed F16 HDR ramp per GPU", "Mblocks/s (4x4) BC6HS", "f16s", 128, 16, "bc6h_encode_kernel<true,false>"),
    "ETC2_RGBA": ("EncodeETC2RGBA Options=default 4096x4096 synthetic RGBA8 (1048576 blocks) per GPU", "Mblocks/s (4x4) ETC2 RGBA", "rgba8", 64, 16, "etc_encode_kernel<2,false,false>"),
    "ETC2": ("EncodeETC2 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) ETC2 RGB", "rgba8", 64, 8, "etc_encode_kernel<1,false,false>"),
    "ETC1": ("EncodeETC1 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) ETC1", "rgba8", 64, 8, "etc_encode_kernel<0,false,false>"),
    "ETC2_ALPHA": ("EncodeETC2Alpha 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) EAC alpha", "rgba8", 64, 8, "eac_encode_kernel<0>"),
    "BC1": ("EncodeBC1 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC1", "rgba8", 64, 8, "s3tc_encode_kernel<BC1>"),
    "BC2": ("EncodeBC2 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC2", "rgba8", 64, 16, "s3tc_encode_kernel<BC2>"),
    "BC3": ("EncodeBC3 Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC3", "rgba8", 64, 16, "s3tc_encode_kernel<BC3>"),
    "BC4U": ("EncodeBC4U Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC4U", "rgba8", 64, 8, "s3tc_encode_kernel<BC4U>"),
    "BC5U": ("EncodeBC5U Options=default 4096x4096 synthetic RGBA8 per GPU", "Mblocks/s (4x4) BC5U", "rgba8", 64, 16, "s3tc_encode_kernel<BC5U>"),
}
FORMAT = "BC7"
WORKLOAD, METRIC, INPUT_KIND, IN_BYTES, OUT_BYTES, KERNEL_NAME = FORMAT_CONFIGS[FORMAT]
ALGO_BYTES_PER_BLOCK = IN_BYTES + OUT_BYTES


def select_format(name):
    global FORMAT, WORKLOAD, METRIC, INPUT_KIND, IN_BYTES, OUT_BYTES, KERNEL_NAME, ALGO_BYTES_PER_BLOCK
    FORMAT = name
    WORKLOAD, METRIC, INPUT_KIND, IN_BYTES, OUT_BYTES, KERNEL_NAME = FORMAT_CONFIGS[name]
    ALGO_BYTES_PER_BLOCK = IN_BYTES + OUT_BYTES


def synthetic_blocks(seed_offset=0):
    from convectionkernels_b200 import synth
    if INPUT_KIND == "f16":
        return synth.image_to_blocks(synth.hdr_ramp_f16(SIDE, SIDE, seed=99 + seed_offset))
    if INPUT_KIND == "f16s":
        return synth.image_to_blocks(synth.hdr_ramp_f16(SIDE, SIDE, seed=99 + seed_offset, signed=True))
    return synth.image_to_blocks(synth.mixed_rgba8(SIDE, SIDE, seed=1234 + seed_offset))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profile_constants():
    """dram traffic per launch and issue-slot utilisation from the committed ncu capture (profiles/), if any."""
    try:
        name = "bc7_kernel_ncu_summary.json" if FORMAT == "BC7" else FORMAT.lower() + "_kernel_ncu_summary.json"
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
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

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref = the unmodified sources compiled by
    oracle/Makefile), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.loader import Reference
    R = Reference()
    threads = R.hardware_threads()
    sample_blocks = 131072 if FORMAT in ("BC7", "BC6HU", "BC6HS") else 524288
    blocks = synthetic_blocks()[:sample_blocks]
    opt, plan = R.default_options(), (R.plan_from_quality(100) if FORMAT == "BC7" else None)
    for _ in range(args.warmup):
        R.encode(FORMAT, blocks[:16384], opt, plan, threads=0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        R.encode(FORMAT, blocks, opt, plan, threads=0)
    dt = time.perf_counter() - t0
    v = sample_blocks * args.steps / dt / 1e6
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mblocks/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "first %d blocks of the texture per step" % sample_blocks},
        "cpu_baseline": {"value": v, "unit": "Mblocks/s", "cores": threads, "kind": "reference", "sample": "%d blocks x %d steps, %d threads" % (sample_blocks, args.steps, threads)},
        "e2e": {"value": v, "unit": "Mblocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def sharded_texture(args):
    """BASELINE.json configs[4]: EncodeBC7 quality 100 on ONE 8192x8192 RGBA8 texture (4 194 304 blocks) that lives on
    rank 0; per step the 8-block-group ranges are scattered to the ranks over NCCL, encoded, and the encoded ranges gathered
    on rank 0 (strong scaling: total work fixed).  Device-timed, max over ranks; prints one JSON line."""
    import torch
    import torch.distributed as dist
    from convectionkernels_b200 import api, synth, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    api.init(local_rank)

    side = 8192
    n_blocks = (side // 4) * (side // 4)
    d_all = None
    if rank == 0:
        img = np.concatenate([np.concatenate([synth.mixed_rgba8(SIDE, SIDE, seed=1234 + 2 * r + c) for c in range(2)], axis=1) for r in range(2)], axis=0)
        d_all = torch.from_numpy(synth.image_to_blocks(img).reshape(-1)).to(dev)
    opt, plan = api.Options(), api.BC7EncodingPlan()
    api.ConfigureBC7EncodingPlanFromQuality(plan, 100)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        if distributed:
            local = sharding.scatter_blocks(d_all, n_blocks, IN_BYTES, src=0, device=dev)
        else:
            local = d_all
        enc = api.encode("BC7", local.reshape(-1, IN_BYTES), opt, plan)
        if distributed:
            return sharding.gather_encoded(enc, n_blocks, OUT_BYTES, dst=0)
        return enc

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = api.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        e0.record()
        out = step()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    barrier()
    launches = api.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([total], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    if rank == 0:
        # size-independent check at full size: the sharded result equals a single-GPU encode of sampled group ranges
        ok = True
        full = out.reshape(n_blocks, OUT_BYTES)
        for first in (0, n_blocks // 2 - 4096, n_blocks - 8192):
            ref = api.encode("BC7", d_all.reshape(n_blocks, IN_BYTES)[first:first + 8192].contiguous(), opt, plan)
            ok = ok and bool((ref == full[first:first + 8192]).all())
        print(json.dumps({
            "metric": METRIC, "value": n_blocks * args.steps / (total_ms / 1e3) / 1e6, "unit": "Mblocks/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "EncodeBC7 plan=FromQuality(100) Options=default ONE 8192x8192 synthetic RGBA8 (4194304 blocks) on rank 0, 8-block-group ranges scattered / gathered over NCCL inside the step",
                       "parallelism": "block-range shard x%d" % world, "l2": "256 MiB flush write between timed iterations"},
            "clocks": sampler.summary(), "gpu_launches": int(launches), "sharded_equals_single_gpu_on_sampled_ranges": ok,
        }))
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--format", default="BC7", choices=sorted(FORMAT_CONFIGS))
    ap.add_argument("--workload", default="texture4096", choices=["texture4096", "shard8192"],
                    help="texture4096: every rank encodes its own 4096x4096 texture (weak scaling, the default contract); "
                         "shard8192: BASELINE.json configs[4], ONE 8192x8192 texture on rank 0, block ranges scattered to the ranks, "
                         "encoded and gathered back inside the timed step (strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    select_format(args.format)

    if args.impl == "reference":
        reference_arm(args)
        return
    if args.workload == "shard8192":
        sharded_texture(args)
        return

    import torch
    import torch.distributed as dist
    from convectionkernels_b200 import api, synth, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    api.init(local_rank)

    # synthetic input: each rank owns one whole texture (weak scaling); pinned host copy for the e2e leg
    blocks_np = synthetic_blocks(rank)
    host_in = torch.from_numpy(blocks_np.reshape(-1)).pin_memory()
    host_out = torch.empty(BLOCKS * OUT_BYTES, dtype=torch.uint8).pin_memory()
    d_in = host_in.to(dev)
    d_out = torch.empty((BLOCKS, OUT_BYTES), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    opt = api.Options()
    plan = None
    if FORMAT == "BC7":
        plan = api.BC7EncodingPlan()
        api.ConfigureBC7EncodingPlanFromQuality(plan, 100)
    total_blocks = BLOCKS * world

    def step():
        api.encode(FORMAT, d_in, opt, plan, out=d_out)
        if distributed:
            return sharding.gather_encoded(d_out, total_blocks, OUT_BYTES, dst=0)
        return d_out

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- device-timed leg -------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_events = []
    launches0 = api.launch_count()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ms_total = 0.0
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)                                   # L2 flush between timed iterations (outside the event pair)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record()
        k0.record()
        api.encode(FORMAT, d_in, opt, plan, out=d_out)
        k1.record()
        if distributed:
            sharding.gather_encoded(d_out, total_blocks, OUT_BYTES, dst=0)
        e_end.record()
        torch.cuda.synchronize()
        step_ms_total += e_start.elapsed_time(e_end)
        kernel_events.append(k0.elapsed_time(k1))
    barrier()
    launches = api.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t = torch.tensor([step_ms_total], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = total_blocks * args.steps / (total_ms / 1e3) / 1e6
    kernel_ms = float(np.mean(kernel_events))

    # ---- end-to-end leg: public call with host buffers, copies inside the timed region ----------------------
    host_out_np = host_out.numpy().reshape(BLOCKS, OUT_BYTES)
    host_in_np = host_in.numpy()
    api.encode(FORMAT, host_in_np, opt, plan, out=host_out_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        api.encode(FORMAT, host_in_np, opt, plan, out=host_out_np)       # returns when host_out is complete
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_blocks * args.steps / float(te.item()) / 1e6
    same = bool((host_out_np == d_out.cpu().numpy()).all())

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = ALGO_BYTES_PER_BLOCK * BLOCKS / (kernel_ms / 1e3) / 1e9
        prof = profile_constants()
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": prof.get("dram_bytes_per_launch"), "peak_source": peak_src,
                    "kernel": KERNEL_NAME, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_BLOCK * BLOCKS,
                    "note": "compute-bound path (millions of instructions per block against <= 144 bytes); the binding unit is the FP32 pipe / issue slots, see fma_pipe_frac_ncu, issue_slot_frac_ncu and DESIGN.md",
                    "issue_slot_frac_ncu": prof.get("issue_slot_frac"), "fma_pipe_frac_ncu": prof.get("fma_pipe_frac")}
        # the same issue-slot fraction from THIS run's kernel time: instructions per block (committed ncu capture) x blocks
        # / 32 lanes, over the 4 schedulers x SMs x the SM clock sampled during the timed region
        ipb, clk = prof.get("warp_instructions_per_block"), sampler.summary().get("sm_mhz")
        if ipb and clk:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            roofline["issue_slot_frac_live"] = ipb * BLOCKS / 32.0 / (kernel_ms / 1e3 * sms * 4 * clk * 1e6)
            roofline["instructions_per_block_ncu"] = ipb

        # CPU baseline: the unmodified reference on this box's host cores, bounded sample of the same texture
        cpu = None
        try:
            if world > 1:
                raise RuntimeError("measured at N=1 only (the other ranks share the host cores)")
            from oracle.loader import Reference
            R = Reference()
            threads = R.hardware_threads()
            sample = blocks_np[:CPU_SAMPLE_BLOCKS]
            ob, pb = np.frombuffer(bytes(memoryview(opt)), np.uint8), (np.frombuffer(plan.tobytes(), np.uint8) if plan is not None else None)
            R.encode(FORMAT, sample[:8192], ob, pb, threads=0)
            t0 = time.perf_counter()
            ref_out = R.encode(FORMAT, sample, ob, pb, threads=0)
            dt = time.perf_counter() - t0
            cpu = {"value": CPU_SAMPLE_BLOCKS / dt / 1e6, "unit": "Mblocks/s", "cores": threads, "kind": "reference",
                   "sample": "first %d blocks (1024 rows) of the texture, %d threads, %.1f s" % (CPU_SAMPLE_BLOCKS, threads, dt),
                   "bit_exact_vs_gpu": bool((ref_out == host_out_np[:CPU_SAMPLE_BLOCKS]).all())}
        except Exception as e:
            cpu = {"value": None, "unit": "Mblocks/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}

        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Mblocks/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "blocks_per_gpu": BLOCKS, "parallelism": "block-range shard x%d, NCCL gather of encoded ranges" % world if distributed else "single GPU",
                       "l2": "256 MiB flush write between timed iterations", "plan": "ConfigureBC7EncodingPlanFromQuality(100)" if FORMAT == "BC7" else None, "flags": "Default (BC7_FastIndexing|S3TC_Paranoid)", "format": FORMAT},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "Mblocks/s", "h2d_bytes_per_step": BLOCKS * IN_BYTES, "d2h_bytes_per_step": BLOCKS * OUT_BYTES,
                    "host_equals_device_result": same},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }))

    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
