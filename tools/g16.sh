#!/bin/bash
# round-1f record: full GPU suite, bench (ours + reference arm), ncu launch list, ncu full capture of the fused kernel at 10M
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01g_10m.json 2> gpurun_out/bench_r01g_10m.err; cat gpurun_out/bench_r01g_10m.json; tail -3 gpurun_out/bench_r01g_10m.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01g_ref.json 2>&1; cat gpurun_out/bench_r01g_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01g.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:cells_fused_kernelILi16 -s 2 -c 1 -f -o gpurun_out/prof_fused_r01g_10m python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_fused_10m.log 2>&1
tail -2 gpurun_out/ncu_fused_10m.log
