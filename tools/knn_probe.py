#!/usr/bin/env python
"""kNN timing probe: grid build + exact kNN (ids only, device output) at n seeds, k neighbours."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geogram_b200 import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [21]
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
x = torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=gen)
ctx = Context(0)
ctx.set_points_device(n, x.data_ptr())
for k in ks:
    out = torch.empty((n, k), dtype=torch.int32, device="cuda")
    ctx.knn_device(k, out.data_ptr())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); 
    for _ in range(3):
        ctx.knn_device(k, out.data_ptr())
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(json.dumps({"n": n, "k": k, "ms_knn": ms, "queries_per_s": n / ms * 1e3, "algorithmic_gbs": (76 + 4 * k) * n / ms / 1e6}), flush=True)
