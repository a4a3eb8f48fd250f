timeout 900 python tools/try_fused.py 20000 2>&1 | tail -60
timeout 300 python bench.py --steps 3 --warmup 3 --seeds 2000000 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l); continue
    print('BENCH value %.3e ms/step %.2f phases %s e2e %.3e'%(d['value'],d['ms_per_step'],d['phase_ms_per_step'],d['e2e']['value']))
"
