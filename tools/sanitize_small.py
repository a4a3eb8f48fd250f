"""Small run of every kernel variant for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracles
from geogram_b200 import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
ctx = Context(0)
x = oracles.uniform_seeds(n, 1)
ctx.set_domain_periodic(1.0); ctx.set_points(x)
ctx.compute_cells(want_faces=2, want_energy=True)
m, mg, st = ctx.get_cells(); print("periodic faces+verts+energy", np.bincount(st), m.sum() - 1)
w = 0.02 * n ** (-2.0 / 3.0) * np.random.default_rng(0).random(n)
ctx.set_points(x, w); ctx.compute_cells(); print("laguerre", ctx.get_cells()[0].sum() - 1)
ctx.lloyd(2)
wn, res = ctx.otm_newton(None, w0=w, epsilon=1e-2, max_newton=3); print("otm", res["newton_iters"], res["max_rel_error"])
ctx2 = Context(0)
xc = oracles.clustered_seeds(n, 2)
ctx2.set_points(xc); ctx2.compute_cells(want_faces=True)
s = ctx2.stats(); print("clustered box", np.bincount(ctx2.get_cells()[2]), s["n_pass"][:s["passes"]])
ctx2.compute_cells(mode=2); print("table pipeline", np.bincount(ctx2.get_cells()[2]))
nb, d2 = ctx2.knn(12); print("knn", d2[:, 1].mean())
