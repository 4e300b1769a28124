#!/usr/bin/env python3
"""profiles/<tag>_misc_kernels.csv from an ncu report with several kernels: duration, DRAM bytes, achieved GB/s against the
measured copy bandwidth, top stall.  usage: summarize_misc.py <tag> <rep | raw csv of `ncu -i rep --page raw --csv`>"""
import csv, io, json, os, subprocess, sys
tag, rep = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "msecond": 1e-3, "second": 1}
def val(r, k):
    return float(r[ix[k]].replace(",", "")) * SCALE.get(units[ix[k]], 1)
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
out = os.path.join(ROOT, "profiles", tag + "_misc_kernels.csv")
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none; second launch of each kernel in tools/prof_misc.py; peak = measured copy bandwidth %.1f GB/s\n" % peak)
    f.write("kernel,grid,block,duration_us,dram_read_MB,dram_write_MB,dram_GBps,frac_of_copy_peak,sm_issue_pct,registers,top_stall,top_stall_per_issue\n")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        top = max(stalls, key=lambda h: float(r[ix[h]].replace(",", "") or 0))
        f.write("\"%s\",\"%s\",\"%s\",%.1f,%.1f,%.1f,%.1f,%.3f,%s,%s,%s,%s\n" % (
            r[ix["Kernel Name"]][:70], r[ix["Grid Size"]], r[ix["Block Size"]], t * 1e6, rd / 1e6, wr / 1e6, (rd + wr) / t / 1e9, (rd + wr) / t / 1e9 / peak,
            r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]], r[ix["launch__registers_per_thread"]],
            top.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[ix[top]]))
print(open(out).read())
