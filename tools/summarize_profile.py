#!/usr/bin/env python3
"""Turns gpurun_out/<name>.ncu-rep (+ launches.csv) into the tracked summaries under profiles/.
usage: summarize_profile.py <tag> <rep> <blocks_in_captured_launch> [launches.csv|-] [summary_name]"""
import csv, io, json, os, subprocess, sys

tag, rep, blocks = sys.argv[1], sys.argv[2], int(sys.argv[3])
launches = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None
summary_name = sys.argv[5] if len(sys.argv) > 5 else "bc7"      # profiles/<summary_name>_kernel_ncu_summary.json, read by bench.py
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keep = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__thread_inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"]
with open(os.path.join(OUT, tag + "_ncu_metrics.csv"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on; one launch of %d blocks; values per launch\n" % blocks)
    f.write("metric,unit,value\n")
    for k in keep:
        if k in m:
            f.write("%s,%s,%s\n" % (k, m[k][1], m[k][0].replace(",", "")))


def num(k):
    v, u = m[k]
    x = float(v.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "msecond": 1e-3, "usecond": 1e-6, "second": 1}.get(u, 1)
    return x * scale


dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
sys.path.insert(0, ROOT)
import bench
summary = {
    "source": os.path.basename(rep), "captured_launch_blocks": blocks,
    "kernel_source_hash": bench.kernel_source_hash(summary_name),
    "kernel_source_hash_note": "sha256[:16] over bench.KERNEL_SOURCES[name] at capture time; bench.py only uses the instruction count while it still matches",
    "kernel_time_s_under_ncu": num("gpu__time_duration.sum"),
    "dram_bytes_per_block": dram / blocks,
    "dram_bytes_per_launch": dram / blocks * 1048576,
    "dram_bytes_per_launch_note": "dram__bytes_read.sum + dram__bytes_write.sum of the captured launch scaled to the 1 048 576-block bench launch",
    "issue_slot_frac": float(m["sm__inst_executed.avg.per_cycle_active"][0]) / 4.0,
    "issue_slot_frac_note": "sm__inst_executed.avg.per_cycle_active / 4 warp-instructions per SM cycle",
    "fma_pipe_frac": float(m["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0]) / 100.0 if "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active" in m else None,
    "fma_pipe_frac_note": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active: share of cycles the FP32 FMA pipe is busy (packed FFMA2/FADD2 occupy it for two cycles)",
    "registers_per_thread": int(float(m["launch__registers_per_thread"][0])),
    "warp_instructions_per_block": float(m["smsp__inst_executed.sum"][0].replace(",", "")) / (blocks / 32.0) if "smsp__inst_executed.sum" in m else None,
}
with open(os.path.join(OUT, summary_name + "_kernel_ncu_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)
print(json.dumps(summary, indent=1))

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
tmp = "/tmp/_sass_%s.csv" % tag
open(tmp, "w").write(sass)
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_summary.py"), tmp, "30"], stdout=subprocess.PIPE, text=True).stdout
open(os.path.join(OUT, tag + "_sass_opcode_mix.txt"), "w").write(txt)

if launches:
    lines = [l for l in open(launches) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    ix = {h: i for i, h in enumerate(rows[0])}
    with open(os.path.join(OUT, tag + "_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 40 python bench.py --steps 2 --warmup 1 (cold-cache, serialised: compare shares)\n")
        f.write("id,kernel,block,grid,duration_ns\n")
        for r in rows[1:]:
            if len(r) >= len(rows[0]):
                f.write("%s,\"%s\",\"%s\",\"%s\",%s\n" % (r[ix["ID"]], r[ix["Kernel Name"]][:90], r[ix["Block Size"]], r[ix["Grid Size"]], r[ix["Metric Value"]]))
