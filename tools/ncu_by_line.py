#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS rows) by CUDA source line / region, using the
line table of `nvdisasm -g` on the same cubin.
usage: ncu_by_line.py <source.csv> <nvdisasm -g output> <kernel-substring> [regions: name:lo-hi,...]"""
import csv, re, sys, collections
src_csv, sass, kname = sys.argv[1:4]
regions = []
if len(sys.argv) > 4:
    for r in sys.argv[4].split(","):
        nm, rg = r.split(":"); lo, hi = rg.split("-"); regions.append((nm, int(lo), int(hi)))
# line table
off2line = {}
cur = None; infn = False
for ln in open(sass):
    if ln.startswith("//---") and ".text." in ln:
        infn = kname in ln
        cur = None
        continue
    if not infn: continue
    if "//## File" in ln:
        # nvdisasm -gi: consecutive annotation lines go from the innermost frame outwards;
        # keep the outermost one (the call site in the kernel body)
        m = re.search(r'inlined at "([^"]+)", line (\d+)', ln) or re.search(r'//## File "([^"]+)", line (\d+)', ln)
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = rows[hdr_i + 1:]
base = int(data[0][col["Address"]], 16)
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
keys = ["Instructions Executed", "Thread Instructions Executed", "# Samples", "stall_wait", "stall_short_sb", "stall_long_sb",
        "stall_no_inst", "stall_branch_resolving", "stall_math", "stall_not_selected", "stall_selected", "stall_mio", "stall_lg", "stall_dispatch",
        "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"]
for r in data:
    if len(r) < len(hdr): continue
    off = int(r[col["Address"]], 16) - base
    line = off2line.get(off)
    for k in keys:
        try: v = float(r[col[k]])
        except Exception: v = 0
        agg[line][k] += v
        tot[k] += v
print("TOTAL", {k: int(v) for k, v in tot.items()})
def show(name, c):
    print("%-28s inst %6.2f%%  thr/inst %5.1f  samples %6.2f%%  wait %5.1f short_sb %5.1f long_sb %5.1f no_inst %5.1f br %5.1f math %5.1f notsel %5.1f sel %5.1f | smem wf %.2fx" % (
        name, 100 * c["Instructions Executed"] / tot["Instructions Executed"],
        c["Thread Instructions Executed"] / max(c["Instructions Executed"], 1),
        100 * c["# Samples"] / tot["# Samples"],
        *[100 * c[k] / max(tot["# Samples"], 1) for k in ("stall_wait", "stall_short_sb", "stall_long_sb", "stall_no_inst", "stall_branch_resolving", "stall_math", "stall_not_selected", "stall_selected")],
        c["L1 Wavefronts Shared"] / max(c["L1 Wavefronts Shared Ideal"], 1)))
if regions:
    for nm, lo, hi in regions:
        c = collections.Counter()
        for line, cc in agg.items():
            if line and line[0] == "fused.cu" and lo <= line[1] <= hi:
                c.update(cc)
        show(nm, c)
    c = collections.Counter()
    for line, cc in agg.items():
        if not line or line[0] != "fused.cu":
            c.update(cc)
    show("(other files / none)", c)
else:
    for line, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:60]:
        show(str(line), c)
