#!/bin/bash
# ncu --set full capture (with source counters) of the team kernel: launch 1 (raw seeds) and launch 5
# (relaxed) of a Lloyd run.  usage: tools/ncu_team.sh [seeds] [tag]
N=${1:-2000000}
TAG=${2:-r02}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:fused_std -s 0 -c 1 \
    -o gpurun_out/${TAG}_team_iter1 -f python bench.py --seeds $N --steps 1 --warmup 0 --no-cpu --no-extra --no-e2e --mode 3 > gpurun_out/${TAG}_team1.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:fused_std -s 4 -c 1 \
    -o gpurun_out/${TAG}_team_iter5 -f python bench.py --seeds $N --steps 5 --warmup 0 --no-cpu --no-extra --no-e2e --mode 3 > gpurun_out/${TAG}_team5.log 2>&1
ls -la gpurun_out/${TAG}_team*.ncu-rep
