#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py 3000 > gpurun_out/sanitizer_memcheck.log 2>&1; tail -12 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py 2000 > gpurun_out/sanitizer_racecheck.log 2>&1; tail -12 gpurun_out/sanitizer_racecheck.log
