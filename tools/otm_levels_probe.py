#!/usr/bin/env python
"""Multilevel OT solve against the single-level solve (geogram_b200/otm.py): Newton / CG counts and
wall time per level.  usage: tools/otm_levels_probe.py [n] [ratio] [amplitude of the target density 1 + a sin(2 pi x)]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geogram_b200 import Context  # noqa: E402
from geogram_b200.otm import optimize_levels  # noqa: E402
from geogram_b200.ordering import brio_order  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ratio = float(sys.argv[2]) if len(sys.argv) > 2 else 0.125
rng = np.random.default_rng(7)
xyz = rng.random((n, 3))
ctx = Context(0)
ctx.set_domain_box([0, 0, 0, 1, 1, 1])
ctx.set_points(xyz)
xyz = ctx.lloyd(5)                     # a sampling of the unit cube, in random (BRIO-like) order
amp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
nu = 1.0 + amp * np.sin(2 * np.pi * xyz[:, 0])
nu /= nu.sum()
order, levels = brio_order(xyz, threshold=300, ratio=ratio)     # GEO::compute_BRIO_order's intervals
xyz = np.ascontiguousarray(xyz[order])
nu = nu[order]
ctx.set_points(xyz)
t0 = time.perf_counter()
w1, r1 = ctx.otm_newton(nu=nu, epsilon=1e-3)
t1 = time.perf_counter() - t0
print("single level: %.3f s" % t1, {k: r1[k] for k in ("newton_iters", "cg_iters", "cell_evals", "converged", "max_rel_error")})
t0 = time.perf_counter()
w2, rs = optimize_levels(ctx, xyz, levels, nu=nu, epsilon=1e-3, log=print)
t2 = time.perf_counter() - t0
for r in rs:
    print("  level %9d:" % r["points"], {k: r[k] for k in ("newton_iters", "cg_iters", "cell_evals", "converged", "max_rel_error", "halvings")})
d = (w2 - w2.mean()) - (w1 - w1.mean())
print("multilevel: %.3f s, levels %s; max |dw| %.3e of range %.3e" % (t2, levels, np.abs(d).max(), np.ptp(w1)))
