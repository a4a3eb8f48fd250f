"""oracle/otm_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of semi-discrete optimal transport by damped Newton on the Laguerre weights.

PARITY UNPINNED against the reference's own solver: OptimalTransportMap3d lives in
BrunoLevy/exploragram, which is not under /root/reference and has no pinned version (SURVEY.md
8(c)); its call sites are src/examples/exploragram/compute_OTM/main.cpp:284-297.  This file
restates the published algorithm instead:
  * B. Levy, "A numerical algorithm for L2 semi-discrete optimal transport in 3D" (2015):
    gradient g_i = vol(Lag_i) - nu_i, Hessian H_ij = A_ij / (2 |p_i - p_j|);
  * Kitagawa, Merigot, Thibert (2019), Algorithm 1: damping so that every cell keeps
    eps0 = 1/2 min(min vol_i(w0), min nu_i) and the error drops by (1 - alpha/2).
Cells, volumes and facet areas come from the oracle port (oracle/voro_oracle.cpp), which IS
pinned to the unmodified reference (tests/test_oracle.py).  Linear systems: sparse direct solve
with one pinned unknown (the Laplacian's null space is the constants).

Only tests/ may import this.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

TRANSLATION = None


def _translations():
    """instance -> translation (delaunay/periodic.cpp:44-86 with reorder_instances)"""
    global TRANSLATION
    if TRANSLATION is None:
        T = np.zeros((27, 3))
        for tx in (-1, 0, 1):
            for ty in (-1, 0, 1):
                for tz in (-1, 0, 1):
                    i = (tz + 1) + 3 * (ty + 1) + 9 * (tx + 1)
                    inst = 0 if i == 13 else (i + 1 if i < 13 else i)
                    T[inst] = (tx, ty, tz)
        TRANSLATION = T
    return TRANSLATION


def laplacian(xyz, R, periodic, period, box_id):
    """L_ii = sum_j c_ij, L_ij = -c_ij, c_ij = A_ij / (2 |p_i - p_j|) from the face lists of R"""
    n = xyz.shape[0]
    rows, cols, vals = [], [], []
    T = _translations()
    for i in range(n):
        nf = int(R.nfaces[i])
        ids = R.face_nbr[i, :nf].astype(np.int64)
        ar = R.face_area[i, :nf]
        keep = ids < box_id
        ids, ar = ids[keep], ar[keep]
        j = ids % n
        inst = ids // n
        q = xyz[j] + (T[inst] * period if periodic else 0.0) - xyz[i]
        c = ar / (2.0 * np.sqrt((q * q).sum(axis=1)))
        rows += [i] * (len(j) + 1)
        cols += list(j) + [i]
        vals += list(-c) + [c.sum()]
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))


def otm_newton(port, xyz, nu=None, w0=None, periodic=True, period=1.0, epsilon=0.01, max_newton=50,
               box_id=0xFFFFFFF0):
    n = xyz.shape[0]
    total = period ** 3 if periodic else 1.0
    nu = np.full(n, total / n) if nu is None else np.asarray(nu, np.float64)
    w = np.zeros(n) if w0 is None else np.array(w0, np.float64)

    def cells(wv):
        return port.cells(xyz, w=wv, periodic=periodic, period=period, kcap=2000, areas=True)

    R = cells(w)
    g = nu - R.m
    eps0 = 0.5 * min(R.m.min(), nu.min())
    assert eps0 > 0
    log = []
    it = 0
    while it < max_newton and (np.abs(g) / nu).max() > epsilon:
        L = laplacian(xyz, R, periodic, period, box_id).tocsc()
        # pin unknown 0 (null space = constants)
        d = np.zeros(n)
        d[1:] = spla.spsolve(L[1:, 1:], g[1:])
        d -= d.mean()
        alpha, ok = 1.0, False
        for _ in range(40):
            Rt = cells(w + alpha * d)
            gt = nu - Rt.m
            if Rt.m.min() >= eps0 and np.linalg.norm(gt) <= (1.0 - 0.5 * alpha) * np.linalg.norm(g):
                ok = True
                break
            alpha *= 0.5
        if not ok:
            break
        w, R, g = w + alpha * d, Rt, gt
        log.append((alpha, float(np.linalg.norm(g)), float((np.abs(g) / nu).max())))
        it += 1
    return w, R, log
