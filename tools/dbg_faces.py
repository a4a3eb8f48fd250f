import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracles
from geogram_b200 import Context
n = int(sys.argv[1])
xyz = oracles.uniform_seeds(n, 12345)
ctx = Context(0); ctx.set_domain_periodic(1.0); ctx.set_points(xyz)
xyz = ctx.lloyd(3); ctx.set_points(xyz)
for faces in (False, True, True):
    for en in (False, True):
        ctx.compute_cells(want_faces=faces, want_energy=en)
        s = ctx.stats()
        print("faces", faces, "energy", en, "ms_clip %.1f ms_knn %.1f passes" % (s["ms_clip"], s["ms_knn"]), s["n_pass"][:s["passes"]], s["k_pass"][:s["passes"]])
