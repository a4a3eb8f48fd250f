import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracles
from geogram_b200 import Context
# degenerate input: regular lattice (every Voronoi vertex has 8 cospherical seeds) + duplicates
g = 12
ax = (np.arange(g) + 0.5) / g
xyz = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3).copy()
xyz = np.concatenate([xyz, xyz[:50]])  # exact duplicates
for periodic in (0, 1):
    ctx = Context(0)
    if periodic: ctx.set_domain_periodic(1.0)
    ctx.set_points(xyz); ctx.compute_cells(want_faces=True)
    m, mg, st = ctx.get_cells(); nf, fn, fa = ctx.get_faces()
    print("lattice periodic=%d" % periodic, "sum-1 = %.3e" % (m.sum() - 1), "status", np.bincount(st).nonzero()[0], np.bincount(st)[np.bincount(st) > 0], "nf hist", np.bincount(nf))
    bad = (st & ~np.uint8(66)) != 0
    print("  other flags:", np.bincount(st[bad]) .nonzero()[0], "n=", bad.sum(), ctx.stats())
