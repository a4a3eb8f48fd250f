set -x
ncu --set full --clock-control none --import-source on -k regex:cells_knn_kernel -s 2 -c 1 -o gpurun_out/prof_clip_r01a python bench.py --steps 1 --warmup 1 --seeds 1000000 --no-cpu > gpurun_out/ncu_clip.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 2 -c 1 -o gpurun_out/prof_knn_r01a python bench.py --steps 1 --warmup 1 --seeds 1000000 --no-cpu > gpurun_out/ncu_knn.log 2>&1
ls -la gpurun_out
