#!/bin/bash
for r in 2 3 4 8; do echo "ring=$r"; B200VOR_WLOC_RING=$r timeout 300 python tools/config_probe.py otm 10000000 5 1e-3 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('wall %.2f s clip %.0f ms newton %d cg %d err %.3e'%(d['wall_s_otm_newton'],d['ms_clip'],d['newton_iters'],d['cg_iters'],d['final_error']))"; done
