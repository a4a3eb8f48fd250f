timeout 600 python tools/try_fused.py 20000 2>&1 | tail -60
