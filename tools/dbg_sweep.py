import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracles
from geogram_b200 import Context
n = 100000
xyz = oracles.clustered_seeds(n, 4321)
ref = oracles.Reference()
A = ref.cells_delaunay(xyz, periodic=False, areas=True)
ctx = Context(0)
ctx.set_points(xyz)
ctx.compute_cells(mode=3, want_faces=True)
s = ctx.stats()
m, mg, st = ctx.get_cells()
nf, fn, fa = ctx.get_faces()
print("passes", s["n_pass"][:s["passes"]], s["k_pass"][:s["passes"]])
print("status hist", dict(zip(*np.unique(st, return_counts=True))))
ok = st == 0
rel = np.abs(m - A.m) / A.m
print("ok cells", ok.sum(), "of which volume wrong (>1e-6):", int((rel[ok] > 1e-6).sum()), "max rel", rel[ok].max())
bad = np.nonzero(ok & (rel > 1e-6))[0][:5]
for i in bad:
    print(i, "m", m[i], "ref", A.m[i], "nf", nf[i], "ref nf", A.nfaces[i])
print("sum m", m.sum(), "sum ok-wrong excess", (m - A.m)[ok & (rel > 1e-6)].sum())
