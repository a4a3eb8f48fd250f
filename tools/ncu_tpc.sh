#!/bin/bash
# ncu --set full capture (with source counters) of the thread-per-cell kernel: launch 1 (raw seeds)
# and launch 5 (relaxed) of a Lloyd run.  usage: tools/ncu_tpc.sh [seeds] [tag]
N=${1:-2000000}
TAG=${2:-r02b}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:cells_tpc_kernel -s 0 -c 1 \
    -o gpurun_out/${TAG}_tpc_iter1 -f python bench.py --seeds $N --steps 1 --warmup 0 --no-cpu --no-extra --no-e2e --mode 4 > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:cells_tpc_kernel -s 4 -c 1 \
    -o gpurun_out/${TAG}_tpc_iter5 -f python bench.py --seeds $N --steps 5 --warmup 0 --no-cpu --no-extra --no-e2e --mode 4 > gpurun_out/${TAG}_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
