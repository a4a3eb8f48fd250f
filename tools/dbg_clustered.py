import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracles
from geogram_b200 import Context
n = 100000
xyz = oracles.clustered_seeds(n, 4321)
ref = oracles.Reference()
A = ref.cells_delaunay(xyz, areas=True)
ctx = Context(0); ctx.set_points(xyz); ctx.compute_cells(want_faces=True)
m, mg, st = ctx.get_cells()
print("status", np.bincount(st), ctx.stats()["n_pass"])
rel = np.abs(m - A.m) / A.m
o = np.argsort(-rel)[:8]
for i in o:
    print(i, "rel %.2e m %.3e nf %d cen_err %.2e size %.2e" % (rel[i], A.m[i], A.nfaces[i], np.abs(mg[i]/m[i] - A.mg[i]/A.m[i]).max(), np.cbrt(A.m[i])))
P = oracles.Port().cells(xyz, kcap=4000)
relp = np.abs(P.m - A.m) / A.m
print("port vs ref max rel", relp.max(), "gpu vs port", (np.abs(m - P.m) / P.m).max())
print("sum", m.sum() - 1, A.m.sum() - 1)
