#!/usr/bin/env python
"""ncu --csv launch list (tools/ncu_inst_by_iter.sh) -> per-iteration instruction counts (JSON).
The bench launches the first-pass cell kernel once per Lloyd iteration; the smaller re-launches
of the flagged cells (other kernel instances) are added to the iteration they belong to."""
import csv
import json
import sys

path, K = sys.argv[1], int(sys.argv[2])
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
launches = {}
for r in csv.DictReader(lines):
    lid = int(r["ID"])
    d = launches.setdefault(lid, {"name": r["Kernel Name"], "grid": r.get("Grid Size", "")})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    d["unit_" + r["Metric Name"]] = r["Metric Unit"]
order = sorted(launches)
iters = []
first = launches[order[0]] if order else None
for lid in order:
    L = launches[lid]
    if L["grid"] == first["grid"]:   # a full-size launch opens an iteration; re-launches are smaller
        iters.append({"grid": L["grid"], "kernel": L["name"], "warp_inst": 0.0, "thread_inst": 0.0,
                      "ns": 0.0, "dram_bytes": 0.0, "launches": 0})
    it = iters[-1]
    it["warp_inst"] += L["smsp__inst_executed.sum"]
    it["thread_inst"] += L["smsp__thread_inst_executed.sum"]
    it["ns"] += L["gpu__time_duration.sum"] * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[L["unit_gpu__time_duration.sum"]]
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        it["dram_bytes"] += L[m] * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[L["unit_" + m]]
    it["launches"] += 1
iters = iters[:K]
out = {"seeds": int(sys.argv[3]) if len(sys.argv) > 3 else 10000000,
       "source": "profiles/r02_inst_by_iter.csv (ncu --metrics smsp__inst_executed.sum,... --clock-control none of "
                 "`bench.py --steps %d --warmup 0`)" % K,
       "iters": iters,
       "dram_bytes_per_launch": sum(i["dram_bytes"] for i in iters) / max(len(iters), 1)}
print(json.dumps(out, indent=1))
