#!/bin/bash
run() {
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu "$@" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.3e ms/step %.2f clip %.2f passes %s'%(d['value'],d['ms_per_step'],d['phase_ms_per_step']['ms_clip'],d['config']['cells_per_pass']))
"
}
for ppc in 5 6 7 8 10 12; do echo -n "ppc=$ppc : "; run --ppc $ppc; done
for t in 30 35 40 45 50; do echo -n "target=$t : "; B200VOR_TARGET=$t run; done
