#!/bin/bash
# state check after re-entry: GPU tests, bench (ours + reference arm), ncu source capture at 2M seeds
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01e_10m.json 2> gpurun_out/bench_r01e_10m.err; cat gpurun_out/bench_r01e_10m.json; tail -3 gpurun_out/bench_r01e_10m.err
bash tools/ncu_fused.sh 2000000
