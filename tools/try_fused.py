"""Dev probe: fused kernel (mode 1) against the legacy pipeline (mode 2) and the CPU oracle."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracles
from geogram_b200 import Context

def run(xyz, w, periodic, mode, faces=True):
    ctx = Context(0)
    if periodic:
        ctx.set_domain_periodic(1.0)
    ctx.set_points(xyz, w)
    ctx.compute_cells(want_faces=faces, mode=mode)
    m, mg, st = ctx.get_cells()
    nf, fn, fa = ctx.get_faces() if faces else (None, None, None)
    s = ctx.stats()
    ctx.close()
    return m, mg, st, nf, fn, fa, s

def compare(name, xyz, w=None, periodic=False):
    A = run(xyz, w, periodic, 1)
    B = run(xyz, w, periodic, 2)
    n = xyz.shape[0]
    stA, stB = A[2], B[2]
    print("== %s n=%d" % (name, n))
    print("   fused  status hist:", {int(k): int(v) for k, v in zip(*np.unique(stA, return_counts=True))}, "ms_clip %.3f" % A[6]["ms_clip"], "P,C,V,I", A[6]["sum_planes"]/n, A[6]["sum_tests"]/n, A[6]["sum_created"]/n, A[6]["sum_tets"]/n)
    print("   legacy status hist:", {int(k): int(v) for k, v in zip(*np.unique(stB, return_counts=True))}, "ms_knn %.3f ms_clip %.3f" % (B[6]["ms_knn"], B[6]["ms_clip"]), "P,C,V,I", B[6]["sum_planes"]/n, B[6]["sum_tests"]/n, B[6]["sum_created"]/n, B[6]["sum_tets"]/n)
    ok = (stA == 0) & (stB == 0)
    col = np.arange(64)[None, :]
    bad = (A[3] != B[3]) | ((A[4] != B[4]) & (col < A[3][:, None])).any(axis=1)
    print("   cells ok in both: %d ; face-list mismatches among them: %d" % (ok.sum(), (bad & ok).sum()))
    if (bad & ok).sum():
        i = np.nonzero(bad & ok)[0][0]
        print("   first bad cell", i, "fused", A[4][i, :A[3][i]], "legacy", B[4][i, :B[3][i]])
    rel = np.abs(A[0][ok] - B[0][ok]) / B[0][ok]
    print("   max rel vol diff %.3e ; max mg diff %.3e ; sum m fused %.12f legacy %.12f" % (rel.max() if rel.size else 0, np.abs(A[1][ok] - B[1][ok]).max() if ok.sum() else 0, A[0].sum(), B[0].sum()))
    if A[5] is not None:
        mask = (col < A[3][:, None]) & ok[:, None] & ~bad[:, None]
        print("   max area diff %.3e" % np.abs(A[5][mask] - B[5][mask]).max())

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    xyz = oracles.uniform_seeds(n, 4242)
    compare("uniform box", xyz)
    compare("uniform periodic", xyz, periodic=True)
    h = n ** (-1.0 / 3)
    w = 0.02 * h * h * np.random.default_rng(7).random(n)
    compare("laguerre periodic", xyz, w=w, periodic=True)
    compare("laguerre box", xyz, w=w)
    compare("clustered box", oracles.clustered_seeds(n, 99))
    big = oracles.uniform_seeds(1000000, 1)
    compare("uniform periodic 1M", big, periodic=True)
