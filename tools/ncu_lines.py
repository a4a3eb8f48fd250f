#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu source-page CSV.
usage: ncu_lines.py <source.csv> <nvdisasm -gi output> <kernel-substring> [lo-hi] [warps]"""
import csv, re, sys, collections
src, sass, kname = sys.argv[1:4]
lo, hi = (0, 10**9)
if len(sys.argv) > 4:
    lo, hi = [int(x) for x in sys.argv[4].split("-")]
warps = float(sys.argv[5]) if len(sys.argv) > 5 else 1e6
off2line = {}; cur = None; infn = False
for ln in open(sass):
    if ln.startswith("//---") and ".text." in ln:
        infn = kname in ln; cur = None; continue
    if not infn: continue
    if "//## File" in ln:
        m = re.search(r'inlined at "([^"]+)", line (\d+)', ln) or re.search(r'//## File "([^"]+)", line (\d+)', ln)
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]; col = {n: i for i, n in enumerate(hdr)}
data = rows[h + 1:]; base = int(data[0][col["Address"]], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0]); tot = 0; ts = 0
for r in data:
    if len(r) < len(hdr): continue
    line = off2line.get(int(r[col["Address"]], 16) - base)
    ie = float(r[col["Instructions Executed"]]); te = float(r[col["Thread Instructions Executed"]]); s = float(r[col["# Samples"]])
    a = agg[line]; a[0] += ie; a[1] += te; a[2] += s; a[3] += 1; tot += ie; ts += s
print("static", len(data), "executed/warp %.0f" % (tot / warps))
for line, a in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))[1]):
    if line and lo <= line[1] <= hi:
        print("%5d inst %5.2f%% /warp %6.0f thr %4.1f samples %5.2f%% static %d" % (line[1], 100 * a[0] / tot, a[0] / warps, a[1] / max(a[0], 1), 100 * a[2] / ts, a[3]))
