#!/bin/bash
# clustered workload (10 M seeds, one GPU) with the current library + the clustered parity tests
timeout 300 python bench.py --workload clustered --seeds ${SEEDS:-10000000} --steps 2 --warmup 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', round(d['ms_per_step'],1),'ms', d['status']['ok'], d['sum_m_minus_1'], d['config']['cells_per_pass_rank0'])"
timeout 400 python -m pytest tests/test_gpu_parity.py -q -k "clustered or two_level or fringe or fuzz" 2>&1 | tail -2
