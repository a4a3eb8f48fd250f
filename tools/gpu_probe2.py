import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracles
from geogram_b200 import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
xyz = oracles.uniform_seeds(n)
ctx2 = Context(0); ctx2.set_domain_periodic(1.0); ctx2.set_points(xyz); ctx2.compute_cells(want_faces=True)
m, mg, st = ctx2.get_cells(); print(m.sum() - 1, np.bincount(st))
