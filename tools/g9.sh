set -x
python - <<'PY'
import torch,time
a=torch.empty(240000000,dtype=torch.uint8).pin_memory(); d=torch.empty(240000000,dtype=torch.uint8,device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(a,non_blocking=True); torch.cuda.synchronize(); h2d=time.perf_counter()-t
    t=time.perf_counter(); a.copy_(d,non_blocking=True); torch.cuda.synchronize(); d2h=time.perf_counter()-t
print("PCIe pinned 240MB: H2D %.1f ms (%.1f GB/s)  D2H %.1f ms (%.1f GB/s)"%(h2d*1e3,0.24/h2d,d2h*1e3,0.24/d2h))
PY
timeout 600 python -m pytest tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -5
nproc
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01c_10m.json 2> gpurun_out/bench_r01c_10m.err; cat gpurun_out/bench_r01c_10m.json; tail -3 gpurun_out/bench_r01c_10m.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01c_ref.json 2>&1; cat gpurun_out/bench_r01c_ref.json
ncu --set full --clock-control none --import-source on -k regex:cells_fused_kernel -s 1 -c 1 -o gpurun_out/prof_fused_r01c_10m python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_fused_10m.log 2>&1
tail -2 gpurun_out/ncu_fused_10m.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
