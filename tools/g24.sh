#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "knn" 2>&1 | tail -3
timeout 200 python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np, oracles
from geogram_b200 import Context
n = 2000000
x = oracles.clustered_seeds(n, 12345)
ctx = Context(0); ctx.set_points(x)
t = time.perf_counter(); nb, d2 = ctx.knn(21); dt = time.perf_counter() - t
print("clustered 2M kNN k=21: %.2f s  (%.2e q/s) two_level grid %s ms_knn %.1f" % (dt, n / dt, ctx.stats()["grid_dim"], ctx.stats()["ms_knn"]))
P = oracles.Port()
idx = np.random.default_rng(0).integers(0, n, 20000)
nbq, d2q = P.knn(x, 21, queries=x[idx])
print("d2 bit-equal on 20000 sampled queries:", bool((d2[idx] == d2q).all()), "ids equal:", bool((nb[idx] == nbq).all()))
PY
