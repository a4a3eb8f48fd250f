#!/bin/bash
run() {
  (cd $1 && timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null) | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.3e ms/step %.2f clip %.2f'%(d['value'],d['ms_per_step'],d['phase_ms_per_step']['ms_clip']))
"
}
echo -n "old r01f: "; run old_r01f
echo -n "current : "; run .
echo -n "old r01f: "; run old_r01f
echo -n "current : "; run .
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or laguerre or reference_live or fused_walk or clustered or fuzz" 2>&1 | tail -3
