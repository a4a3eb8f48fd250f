"""A/B of two builds of the library on the same OT problem (B200VOR_LIB selects the build)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracles
from geogram_b200 import Context
n = int(sys.argv[1])
xyz = oracles.uniform_seeds(n, 5)
ctx = Context(0); ctx.set_domain_periodic(1.0); ctx.set_points(xyz)
xyz = ctx.lloyd(3); ctx.set_points(xyz)
nu = 0.5 + np.random.default_rng(1).random(n); nu /= nu.sum()
t = time.perf_counter(); w, r = ctx.otm_newton(nu, epsilon=1e-6, cg_tol=1e-10, max_cg=5000); dt = time.perf_counter() - t
np.save(sys.argv[2], w - w.mean())
print(json.dumps({"wall_s": dt, **r}))
