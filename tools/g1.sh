set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err
cat gpurun_out/bench_10m.json
python bench.py --steps 3 --warmup 3 --seeds 1000000 --no-cpu > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
cat gpurun_out/bench_1m.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --seeds 2000000 --no-cpu > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
