#!/bin/bash
# A/B of the thread-per-cell kernel's launch shapes on one box (development; not a bench value source)
mkdir -p gpurun_out
for bt in ${BTS:-96 192 224}; do
  B200VOR_TPC_BT=$bt python bench.py --steps 5 --warmup 1 --no-cpu --no-extra --no-e2e --mode ${MODE:-4} > gpurun_out/tpc_bt$bt.json 2> gpurun_out/tpc_bt$bt.err
  python -c "import json;d=json.load(open('gpurun_out/tpc_bt$bt.json'));print('bt',$bt,'cells/s %.1fM'%(d['value']/1e6),[round(x,1) for x in d['ms_iter']],d['config']['cells_per_pass'])"
done
