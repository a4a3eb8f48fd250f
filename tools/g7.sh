for mb in 3 4 5; do
echo -n "minblocks=$mb : "
B200VOR_LIB=$PWD/geogram_b200/lib/libb200vor_mb$mb.so timeout 300 python bench.py --steps 3 --warmup 3 --seeds 2000000 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.3e ms/step %.2f passes %s'%(d['value'],d['ms_per_step'],d['config']['cells_per_pass']))
"
done
