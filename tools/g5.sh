timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l); continue
    print('BENCH value %.3e ms/step %.2f phases %s e2e %.3e passes %s'%(d['value'],d['ms_per_step'],d['phase_ms_per_step'],d['e2e']['value'],d['config']['cells_per_pass']))
"
