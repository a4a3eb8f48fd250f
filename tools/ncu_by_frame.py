#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS rows) by the source line of the KERNEL BODY, i.e. the
inline frame that was inlined directly at `callsite_file` (for cells_tpc_kernel: the line of
compute_cell() in tpc_cell.cuh), using the line table of `nvdisasm -gi` on the same cubin.
usage: ncu_by_frame.py <source.csv> <nvdisasm -gi output> <kernel-substring> <callsite_file> [name:lo-hi,...]"""
import collections
import csv
import re
import sys

src_csv, sass, kname, callsite = sys.argv[1:5]
regions = []
if len(sys.argv) > 5:
    for r in sys.argv[5].split(","):
        nm, rg = r.split(":")
        lo, hi = rg.split("-")
        regions.append((nm, int(lo), int(hi)))
off2line = {}
stack = []
infn = False
fresh = True
for ln in open(sass):
    if ln.startswith("//---") and ".text." in ln:
        infn = kname in ln
        stack = []
        continue
    if not infn:
        continue
    if "//## File" in ln:
        if fresh:
            stack = []
            fresh = False
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        stack.append((m.group(1).split("/")[-1], int(m.group(2)), (m.group(3) or "").split("/")[-1]))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        fresh = True
        cur = None
        for f, l, at in stack:
            if at == callsite:
                cur = (f, l)
        if cur is None and stack:
            cur = stack[-1][:2]
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) >= len(hdr)]
base = int(data[0][col["Address"]], 16)
keys = ["Instructions Executed", "Thread Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"]
stall_cols = [h for h in hdr if h.startswith("stall_")]
agg = collections.defaultdict(collections.Counter)
tot = collections.Counter()
for r in data:
    off = int(r[col["Address"]], 16) - base
    line = off2line.get(off)
    for k in keys + stall_cols:
        try:
            v = float(r[col[k]])
        except Exception:
            v = 0
        agg[line][k] += v
        tot[k] += v
print("TOTAL inst %.3e thread-inst %.3e samples %d" % (tot["Instructions Executed"], tot["Thread Instructions Executed"], tot["# Samples"]))


def show(name, c):
    top = sorted(((c[k], k) for k in stall_cols), reverse=True)[:4]
    print("%-34s inst %6.2f%%  thr/inst %5.1f  thr-inst %6.2f%%  samples %6.2f%% | %s" % (
        name, 100 * c["Instructions Executed"] / tot["Instructions Executed"],
        c["Thread Instructions Executed"] / max(c["Instructions Executed"], 1),
        100 * c["Thread Instructions Executed"] / tot["Thread Instructions Executed"],
        100 * c["# Samples"] / max(tot["# Samples"], 1),
        " ".join("%s %.1f" % (k[6:], 100 * v / max(tot["# Samples"], 1)) for v, k in top)))


if regions:
    rest = collections.Counter()
    regc = [collections.Counter() for _ in regions]
    for line, cc in agg.items():
        hit = False
        if line:
            for k, (nm, lo, hi) in enumerate(regions):
                if lo <= line[1] <= hi and line[0] != callsite:
                    regc[k].update(cc)
                    hit = True
                    break
        if not hit:
            rest.update(cc)
    for (nm, lo, hi), c in zip(regions, regc):
        show("%s [%d-%d]" % (nm, lo, hi), c)
    show("(rest)", rest)
else:
    for line, c in sorted(agg.items(), key=lambda kv: -kv[1]["Instructions Executed"])[:70]:
        show(str(line), c)
