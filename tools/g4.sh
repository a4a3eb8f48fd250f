ncu --set full --clock-control none --import-source on -k regex:cells_fused_kernel -s 1 -c 1 -o gpurun_out/prof_fused_r01b python bench.py --steps 1 --warmup 1 --seeds 1000000 --no-cpu > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
python bench.py --steps 3 --warmup 3 --seeds 2000000 --no-cpu
