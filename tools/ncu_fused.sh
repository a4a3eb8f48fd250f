#!/bin/bash
# ncu --set full capture of the fused kernel with source correlation (run under gpurun)
SEEDS=${1:-2000000}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cells_fused -s 2 -c 1 \
  -f -o gpurun_out/fused_prof python bench.py --seeds $SEEDS --steps 1 --warmup 2 --no-cpu > gpurun_out/ncu_fused.log 2>&1
tail -3 gpurun_out/ncu_fused.log
ls -la gpurun_out/
