import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geogram_b200 import Context
n = 10000000
host_in = torch.rand((n, 3), dtype=torch.float64).pin_memory()
host_out = torch.empty((n, 3), dtype=torch.float64).pin_memory()
hin, hout = host_in.numpy(), host_out.numpy()
ctx = Context(0); ctx.set_domain_periodic(1.0)
def T(f):
    torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3
for it in range(3):
    a = T(lambda: ctx.set_points(hin))
    b = T(lambda: ctx.lloyd(1, out=hout))
    s = ctx.stats()
    print("set_points %.1f ms  lloyd %.1f ms  (stats grid %.1f knn %.1f clip %.1f passes %s)" % (a, b, s["ms_grid"], s["ms_knn"], s["ms_clip"], s["n_pass"][:s["passes"]]))
    hin[:] = hout
c = T(lambda: ctx.lloyd(1, want_xyz=False))
print("lloyd no-out %.1f" % c)
