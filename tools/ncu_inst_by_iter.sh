#!/bin/bash
# Per-launch instruction counts and DRAM bytes of the first-pass cell kernel over the timed Lloyd
# iterations of the DEFAULT bench command (10 M raw uniform seeds, iterations 1..K), for
# bench.py's issue-rate roofline.  Run under gpurun; writes gpurun_out/r02_inst_by_iter.csv and
# .json (copy the .json to profiles/).  `--warmup 0`: launches 1..K of the kernel ARE the timed
# iterations (the same seeds and iterations as a normal run; nothing printed here is a bench value).
set -e
K=${1:-10}
mkdir -p gpurun_out
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:'cells_(tpc|fused)_kernel' --csv --log-file gpurun_out/r02_inst_by_iter.csv \
    python bench.py --steps $K --warmup 0 --no-cpu --no-extra --no-e2e > gpurun_out/r02_inst_bench.log 2>&1
python tools/ncu_inst_parse.py gpurun_out/r02_inst_by_iter.csv $K > gpurun_out/r02_inst_by_iter.json
