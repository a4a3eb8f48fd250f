"""development probe: lattice / tiny inputs through the cell pipeline, per mode"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracles
from conftest import Golden
from geogram_b200 import Context
G = Golden("cells_box_lattice_exact_1728.npz")
for mode in (4, 0):
    for faces in (True, False):
        ctx = Context(0)
        ctx.set_domain_box([0, 0, 0, 0.75, 0.75, 0.75])
        ctx.set_points(G.xyz)
        ctx.compute_cells(mode=mode, want_faces=faces)
        s = ctx.stats()
        m, mg, st = ctx.get_cells()
        print("lattice mode", mode, "faces", faces, "status", {int(v): int((st == v).sum()) for v in np.unique(st)},
              "passes", s["n_pass"][:s["passes"]], "exact", s["exact_cells"], "tpc", s["tpc_cells"], "sum m", m.sum(), "max|m-G|", np.abs(m - G.m).max())
        if faces:
            nf, fn, fa = ctx.get_faces()
            print("   nfaces hist", np.bincount(nf), "golden", np.bincount(G.nfaces), "mismatch cells", int(((nf != G.nfaces) | (fn != G.face_nbr).any(axis=1)).sum()))
for n in (1, 2, 5, 40, 3000):
    p = oracles.uniform_seeds(n, n)
    ctx = Context(0)
    ctx.set_points(p)
    ctx.compute_cells(mode=4)
    m, mg, st = ctx.get_cells()
    s = ctx.stats()
    print("n", n, "status", {int(v): int((st == v).sum()) for v in np.unique(st)}, "sum m - 1", m.sum() - 1.0, "passes", s["n_pass"][:s["passes"]], "grid", s["grid_dim"])

# exact predicate on the device vs the host emulation library
import ctypes
lib = ctypes.CDLL(os.path.join(ROOT, "tests", "_build", "libtpc_host.so")) if os.path.exists(os.path.join(ROOT, "tests", "_build", "libtpc_host.so")) else None
rng = np.random.default_rng(1)
cases = []
for it in range(600):
    Pl = rng.standard_normal((4, 4))
    if it % 3 == 0:
        v = np.linalg.solve(Pl[:3, :3], -Pl[:3, 3]); Pl[3, 3] = -(Pl[3, :3] @ v)
    if it % 3 == 1:
        Pl = np.round(Pl * 8) / 8
        if abs(np.linalg.det(Pl[:3, :3])) < 1e-9: continue
        v = np.round(np.linalg.solve(Pl[:3, :3], -Pl[:3, 3]) * 64) / 64
        for k in range(4): Pl[k, 3] = -(Pl[k, :3] @ v)
    cases.append(Pl)
cases = np.array(cases)
ctx = Context(0)
dev = ctx.exact_signs(cases)
if lib is not None:
    host = np.array([lib.tpc_exact_sign(*[np.ascontiguousarray(r).ctypes.data_as(ctypes.c_void_p) for r in c]) for c in cases])
    print("exact sign device vs host: mismatches", int((dev != host).sum()), "of", len(cases), "zeros host", int((host == 0).sum()), "zeros dev", int((dev == 0).sum()))
else:
    print("dev signs hist", np.bincount(dev + 1))
