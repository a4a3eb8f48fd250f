for ppc in 8 10 12 16 20; do for tg in 36 48; do
echo -n "ppc=$ppc target=$tg : "
B200VOR_TARGET=$tg timeout 300 python bench.py --steps 3 --warmup 3 --seeds 2000000 --no-cpu --ppc $ppc 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.3e ms/step %.2f passes %s'%(d['value'],d['ms_per_step'],d['config']['cells_per_pass']))
"
done; done
