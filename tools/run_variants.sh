#!/bin/bash
# on the GPU box: bench every variant library (10M seeds, device-timed), checking the smoke parity first
SEEDS=${SEEDS:-10000000}
for so in geogram_b200/lib/variants/*.so; do
  echo "=== $so"
  B200VOR_LIB=$PWD/$so timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  B200VOR_LIB=$PWD/$so timeout 300 python bench.py --steps 8 --warmup 1 --seeds $SEEDS --no-cpu --no-extra --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.3e ms/step %.2f iters %s'%(d['value'],d['ms_per_step'],[round(x,1) for x in d['ms_iter']]))
"
done
