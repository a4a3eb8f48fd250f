#!/usr/bin/env python
"""Diagnosis: Laguerre cells of a SMOOTH large-displacement weight field (OT towards the density 1 + x)
on the device against the CPU port: status histogram per mode, passes, volumes.
usage: tools/otm_smooth_probe.py [n]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracles  # noqa: E402
import otm_oracle  # noqa: E402
from geogram_b200 import Context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
port = oracles.Port()
ctx = Context(0)
ctx.set_domain_box([0, 0, 0, 1, 1, 1])
ctx.set_points(oracles.uniform_seeds(n, 31))
xyz = ctx.lloyd(5)
nu = 1.0 + xyz[:, 0]
nu /= nu.sum()
t = time.perf_counter()
w_cpu, R, log = otm_oracle.otm_newton(port, xyz, nu, periodic=False, epsilon=1e-3, max_newton=30)
print("n", n, "cpu newton", [(a, round(e, 6)) for a, _, e in log], "%.1f s" % (time.perf_counter() - t),
      "w range %.3e, h^2 %.3e" % (np.ptp(w_cpu), n ** (-2.0 / 3.0)))
names = ["uncertified", "uncertain", "empty", "boxface", "overflow", "duplicate", "degenerate"]
for s in (0.25, 0.5, 1.0):
    w = s * w_cpu
    C = port.cells(xyz, w=w, periodic=False, kcap=2000)
    for mode in (0, 3, 2):
        ctx.set_points(xyz, w)
        try:
            ctx.compute_cells(want_faces=True, mode=mode)
        except Exception as err:
            print("  s", s, "mode", mode, "error", err)
            continue
        m, mg, st = ctx.get_cells()
        nf, fn, fa = ctx.get_faces()
        S = ctx.stats()
        hist = {nm: int(((st >> k) & 1).sum()) for k, nm in enumerate(names) if ((st >> k) & 1).any()}
        bad = np.nonzero(st & ~np.uint8(4))[0]
        ok = st == 0
        print("  s", s, "mode", mode, "flags", hist, "passes", S["n_pass"][: S["passes"]], "k", S["k_pass"][: S["passes"]],
              "max faces", int(nf.max()), "port max faces", int(C.nfaces.max()) if hasattr(C, "nfaces") else None,
              "max |m - port| on OK cells %.2e" % (np.abs(m[ok] - C.m[ok]).max() if ok.any() else -1))
        if len(bad):
            i = bad[:6]
            print("      flagged cells", i.tolist(), "status", st[i].tolist(), "x", np.round(xyz[i, 0], 3).tolist(),
                  "port m*n", np.round(C.m[i] * n, 3).tolist(), "gpu m*n", np.round(m[i] * n, 3).tolist())
ctx.set_points(xyz)
for mx in (50,):
    try:
        w, r = ctx.otm_newton(nu=nu, epsilon=1e-3, max_newton=mx)
        print("gpu newton", {k: r[k] for k in ("newton_iters", "cg_iters", "cell_evals", "converged", "max_rel_error", "last_alpha")})
    except Exception as err:
        print("gpu newton error", err)
