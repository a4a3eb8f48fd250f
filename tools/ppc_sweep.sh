#!/bin/bash
# A/B of the seeds-per-bin target of the grid on the default bench (raw seeds, iterations 1..6)
for p in "$@"; do
  timeout 120 python bench.py --steps 6 --warmup 1 --no-cpu --no-extra --no-e2e --ppc $p 2>/dev/null > gpurun_out/ppc_$p.json
  python -c "
import json
d=json.load(open('gpurun_out/ppc_$p.json')); print('ppc', $p, round(d['value']/1e6,2), [round(x,1) for x in d['ms_iter']], d['config']['grid'])"
done
