timeout 900 python tools/try_fused.py 20000 2>&1 | grep -E "==|status hist|mismatch"
for ppc in 3 4 6 8; do
echo -n "ppc=$ppc : "
B200VOR_BENCH_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 3 --seeds 2000000 --no-cpu --ppc $ppc 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.3e ms/step %.2f e2e %.3e scans/cell %.2f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['flop_model']['grid_scans_per_cell']))
"
done
python tools/e2e_probe.py 2>&1 | grep -E "set_points|lloyd"
