"""Ad-hoc GPU diagnostics (not a test): prints parity numbers against the oracles."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracles
from geogram_b200 import Context

def cmp_cells(name, ctx, O, n):
    m, mg, st = ctx.get_cells()
    nf, fn, fa = ctx.get_faces()
    class G: pass
    g = G(); g.nfaces = nf; g.face_nbr = fn; g.fcap = 64
    mism = oracles.face_mismatches(g, O)
    ok = st == 0
    rel = np.abs(m - O.m) / np.maximum(O.m, 1e-300)
    print(name, "sum(m)-1 = %.3e" % (m.sum() - 1), "status hist", np.bincount(st, minlength=1)[:40].nonzero()[0], np.bincount(st)[np.bincount(st)>0],
          "face mismatches", mism, "max rel vol %.3e" % rel[ok].max(), "max mg diff %.3e" % np.abs(mg - O.mg)[ok].max())
    if mism:
        bad = np.nonzero((nf != O.nfaces) | ((fn[:, :O.fcap] != O.face_nbr) & (np.arange(64)[None,:] < nf[:,None])).any(1))[0]
        for i in bad[:5]:
            print("  cell", i, "st", st[i], "gpu", fn[i, :nf[i]], "ora", O.face_nbr[i, :O.nfaces[i]], m[i], O.m[i])
    print("  stats", {k: v for k, v in ctx.stats().items() if k not in ("reserved",)})

def main():
    P = oracles.Port()
    R = oracles.Reference() if oracles.have_reference() else None
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    xyz = oracles.uniform_seeds(n)
    ctx = Context(0)
    ctx.set_points(xyz)
    nb, d2 = ctx.knn(21)
    nb2, d22 = P.knn(xyz, 21)
    print("knn: d2 bit-equal", (d2 == d22).all(), "idx equal", (nb == nb2).mean())
    if R:
        nb3, d23 = R.knn(xyz, 21)
        print("knn vs reference: d2 bit-equal", (d2 == d23).all(), "idx equal", (nb == nb3).mean())
    q = np.random.default_rng(3).random((1000, 3)) * 1.2 - 0.1
    nbq, d2q = ctx.knn_points(q, 8); nbq2, d2q2 = P.knn(xyz, 8, queries=q)
    print("knn_points: d2 bit-equal", (d2q == d2q2).all(), "idx equal", (nbq == nbq2).mean())
    # box cells
    ctx.compute_cells(want_faces=True)
    O = P.cells(xyz)
    cmp_cells("box", ctx, O, n)
    # periodic
    ctx2 = Context(0); ctx2.set_domain_periodic(1.0); ctx2.set_points(xyz); ctx2.compute_cells(want_faces=True)
    O = P.cells(xyz, periodic=True)
    cmp_cells("periodic", ctx2, O, n)
    if R:
        A = R.cells_delaunay(xyz, periodic=True)
        cmp_cells("periodic vs ref A", ctx2, A, n)
    # weighted periodic
    h = n ** (-1 / 3.0); w = 0.02 * h * h * np.random.default_rng(5).random(n)
    ctx3 = Context(0); ctx3.set_domain_periodic(1.0); ctx3.set_points(xyz, w); ctx3.compute_cells(want_faces=True)
    O = P.cells(xyz, w=w, periodic=True)
    cmp_cells("weighted periodic", ctx3, O, n)
    # lloyd
    x1 = ctx2.lloyd(2)
    x = xyz
    for it in range(2):
        O = P.cells(x, periodic=True); x = P.lloyd_update(x, O.m, O.mg, True, 1.0)
    print("lloyd 2 it: max diff %.3e" % np.abs(x1 - x).max())
    # timing at larger n
    for nn in (1000000,):
        xx = oracles.uniform_seeds(nn, 7)
        c = Context(0); c.set_domain_periodic(1.0)
        for rep in range(2):
            t = time.time(); c.set_points(xx); c.compute_cells(); dt = time.time() - t
            s = c.stats()
            print("n=%d wall %.3f s  grid %.2f knn %.2f clip %.2f ms passes %s %s status unc=%d uncert=%d ovf=%d box=%d" % (
                nn, dt, s["ms_grid"], s["ms_knn"], s["ms_clip"], s["n_pass"][:s["passes"]], s["k_pass"][:s["passes"]],
                s["n_uncertified"], s["n_uncertain"], s["n_overflow"], s["n_boxface"]))
        m, mg, st = c.get_cells(); print("  sum(m)-1 = %.3e" % (m.sum() - 1), "P,C,V,I per cell", s["sum_planes"]/nn, s["sum_tests"]/nn, s["sum_created"]/nn, s["sum_tets"]/nn)

if __name__ == "__main__":
    main()
