#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:cells_fused_kernelILi16 -s 2 -c 1 -f -o gpurun_out/prof_fused_r01f_10m python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_fused_10m.log 2>&1
tail -2 gpurun_out/ncu_fused_10m.log
