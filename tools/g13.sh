#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.3e ms/step %.2f clip %.2f e2e %.3e passes %s'%(d['value'],d['ms_per_step'],d['phase_ms_per_step']['ms_clip'],d['e2e']['value'],d['config']['cells_per_pass']))
"
