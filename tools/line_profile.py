#!/usr/bin/env python3
"""Source-line view of an ncu capture for the header-inlined device code (ncu's own CUDA view only lists the .cu).

Joins `ncu -i rep --page source --csv --print-source sass` (per-SASS-instruction counters) with the line table of the
same cubin (`nvdisasm -g`), by instruction offset inside the kernel.
usage: line_profile.py <sass.csv> <lib.so> <kernel-substring> [topN]
"""
import collections, csv, os, re, subprocess, sys, tempfile

sass_csv, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
dis = []
for f in sorted(os.listdir(tmp)):          # one cubin per translation unit
    dis += subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout.splitlines()

# line table of the kernel: offset -> inline chain [(file, line) innermost first]; a chain persists until the next one
table, chain, fresh, inside = {}, [("?", 0)], True, False
for ln in dis:
    if ln.startswith(".text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
    if m:
        table[int(m.group(1), 16)] = list(chain)
        fresh = True

rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
base = int(body[0][ix["Address"]], 16)
inst, samp, incl, incl_s = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
tot_i = tot_s = 0
for r in body:
    off = int(r[ix["Address"]], 16) - base
    ch = table.get(off, [("?", 0)])
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    # exclusive: innermost frame that is not the arithmetic helper header; inclusive: every distinct frame of the chain
    key = next((c for c in ch if c[0] != "cvtt_common.cuh"), ch[0])
    inst[key] += n
    samp[key] += s
    for c in set(ch):
        if c[0] != "cvtt_common.cuh":
            incl[c] += n
            incl_s[c] += s
    tot_i += n
    tot_s += s
print("kernel %s: %d warp instructions, %d samples, %d SASS lines, %d mapped" % (kern, tot_i, tot_s, len(body), len(table)))
srcs = {}
def text(f, l):
    if f not in srcs:
        for d in ("convectionkernels_b200/csrc", "."):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
            if os.path.exists(p):
                srcs[f] = open(p).read().splitlines()
                break
        else:
            srcs[f] = []
    return srcs[f][l - 1].strip()[:110] if 0 < l <= len(srcs[f]) else ""
print("%-26s %7s %7s  %s" % ("file:line", "inst%", "samp%", "source"))
for key, n in inst.most_common(top):
    print("%-26s %6.2f%% %6.2f%%  %s" % ("%s:%d" % key, 100.0 * n / tot_i, 100.0 * samp[key] / max(tot_s, 1), text(*key)))
print()
print("inclusive (every frame of the inline chain):")
for key, n in incl.most_common(top):
    print("%-26s %6.2f%% %6.2f%%  %s" % ("%s:%d" % key, 100.0 * n / tot_i, 100.0 * incl_s[key] / max(tot_s, 1), text(*key)))
