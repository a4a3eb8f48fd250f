#!/bin/bash
# on the GPU box: the clustered workload (BASELINE configs[3], 10 M seeds) with every variant library
for so in geogram_b200/lib/variants/*.so; do
  echo "=== $so"
  B200VOR_LIB=$PWD/$so timeout 300 python bench.py --workload clustered --seeds ${SEEDS:-10000000} --steps 2 --warmup 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', round(d['ms_per_step'],1),'ms', d['status']['ok'], d['sum_m_minus_1'], d['config']['cells_per_pass_rank0'])"
done
