#!/usr/bin/env python3
"""Summarises an `ncu --page source --csv --print-source sass` dump: executed warp instructions by opcode, hottest
address ranges, stall samples.  usage: sass_summary.py sass.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix["Instructions Executed"]]) for r in body)
byop = collections.Counter(); samp = collections.Counter()
for r in body:
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("LD", "ST", "MUFU", "F2", "I2")) else "")
    byop[op] += int(r[ix["Instructions Executed"]])
    samp[op] += int(r[ix["# Samples"]])
print("total warp instructions", tot, " SASS lines", len(body))
ts = sum(samp.values())
for op, c in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print("%-22s %14d %6.2f%%   samples %5.2f%%" % (op, c, 100.0 * c / tot, 100.0 * samp[op] / max(ts, 1)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print("stall samples:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
