"""GPU parity of cells restricted to a convex polyhedral domain (b200vor_set_domain_planes): the first
slice of the mesh-restricted cells of the reference (voronoi/generic_RVD.h:1464-1597 clips every cell by
the tetrahedra of a domain mesh; one convex region = ConvexCell::clip_by_plane by its planes,
voronoi/convex_cell.cpp:467-516).  Oracle: the unmodified reference (oracle B with the domain planes
clipped by its own ConvexCell), and this repo's port."""
import numpy as np
import pytest

import oracles
from test_oracle import _convex_domain

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Context():
    from geogram_b200 import Context as C
    return C


@pytest.mark.parametrize("weighted", [False, True])
def test_cells_in_convex_domain_vs_reference(Context, reference, weighted):
    n = 60000
    xyz = oracles.uniform_seeds(n, 33)
    dom = _convex_domain()
    w = None
    if weighted:
        h = n ** (-1.0 / 3.0)
        w = 0.1 * h * h * np.random.default_rng(4).random(n)
    A = reference.cells_knn(xyz, w=w, domain=dom, areas=True)
    ctx = Context(0)
    ctx.set_domain_planes(dom)
    ctx.set_points(xyz, w)
    ctx.compute_cells(want_faces=True)
    m, mg, st = ctx.get_cells()
    nf, fn, fa = ctx.get_faces()
    # a seed outside the domain owns nothing: EMPTY; nothing else may be flagged
    assert ((st & ~np.uint8(4)) == 0).all(), dict(zip(*np.unique(st, return_counts=True)))
    live = A.m > 0
    assert ((st == 0) == live).all() or (np.abs(A.m[(st == 0) != live]).max() < 1e-12)
    assert np.abs(m - A.m).max() <= 1e-9 * A.m.max()
    assert np.abs(mg - A.mg).max() <= 1e-9 * A.m.max()
    assert abs(m.sum() - A.m.sum()) < 1e-10
    ok = (st == 0) & (A.nfaces <= 64)
    assert (nf[ok] == A.nfaces[ok]).all()
    col = np.arange(64)[None, :]
    mask = (col < nf[:, None]) & ok[:, None]
    assert (fn[mask] == A.face_nbr[mask]).all()            # neighbour ids, box ids and DOMAIN ids, same order
    assert np.abs(fa[mask] - A.face_area[mask]).max() < 1e-7
    ids = fn[mask]
    assert ((ids >= oracles.DOMAIN_ID) & (ids < oracles.DOMAIN_ID + 4)).sum() > 1000
    ctx.close()


def test_lloyd_in_convex_domain_and_api_checks(Context, port):
    from geogram_b200 import B200VorError
    n = 30000
    dom = _convex_domain()
    xyz = oracles.uniform_seeds(n, 35)
    inside = ((xyz @ dom[:, :3].T + dom[:, 3]) >= 0).all(axis=1)
    xyz = xyz[inside]
    ctx = Context(0)
    ctx.set_domain_planes(dom)
    ctx.set_points(xyz)
    x1 = ctx.lloyd(2)
    # the relaxed seeds stay inside the (convex) domain; one step equals the oracle's
    assert ((x1 @ dom[:, :3].T + dom[:, 3]) >= -1e-12).all()
    O = port.cells(xyz, domain=dom)
    x0 = port.lloyd_update(xyz, O.m, O.mg)
    ctx.set_points(xyz)
    xa = ctx.lloyd(1)
    assert np.abs(xa - x0).max() < 1e-9
    with pytest.raises(B200VorError):
        ctx.set_domain_periodic(1.0)                      # domain planes are for box domains
    ctx.set_domain_planes(None)
    ctx.set_domain_periodic(1.0)
    with pytest.raises(B200VorError):
        ctx.set_domain_planes(dom)
    ctx.close()


def _kuhn_tets(i, j, k, nx):
    """the six tetrahedra of the Kuhn subdivision of the unit-lattice cube (i, j, k), vertex ids in an
    (nx+1)^3 lattice"""
    import itertools
    def vid(a, b, c):
        return a + (nx + 1) * (b + (nx + 1) * c)
    out = []
    for perm in itertools.permutations(range(3)):
        p = [i, j, k]
        t = [vid(*p)]
        for ax in perm:
            p[ax] += 1
            t.append(vid(*p))
        out.append(t)
    return out


def _lattice_mesh(nx, keep):
    g = np.arange(nx + 1) / float(nx)
    V = np.array([[g[a], g[b], g[c]] for c in range(nx + 1) for b in range(nx + 1) for a in range(nx + 1)])
    T = []
    for k in range(nx):
        for j in range(nx):
            for i in range(nx):
                if keep(i, j, k):
                    T += _kuhn_tets(i, j, k, nx)
    return V, np.array(T, np.uint32)


def test_tet_mesh_domain_cube_equals_box_domain(Context):
    """A cube cut into 6 x 8 tetrahedra is the cube: cells integrated over the tetrahedra and summed
    (b200vor_set_domain_tets) must equal the cells of the plain box domain."""
    n = 40000
    xyz = oracles.uniform_seeds(n, 41)
    a = Context(0)
    a.set_points(xyz)
    a.compute_cells()
    m0, mg0, st0 = a.get_cells()
    V, T = _lattice_mesh(2, lambda i, j, k: True)
    b = Context(0)
    b.set_domain_tets(V, T)
    b.set_points(xyz)
    b.compute_cells()
    m, mg, st = b.get_cells()
    assert (st == 0).all(), dict(zip(*np.unique(st, return_counts=True)))
    assert np.abs(m - m0).max() <= 1e-12 and np.abs(mg - mg0).max() <= 1e-12
    assert abs(m.sum() - 1.0) < 1e-11
    x1 = b.lloyd(1)
    a.set_points(xyz)
    x0 = a.lloyd(1)
    assert np.abs(x1 - x0).max() < 1e-10


def test_tet_mesh_domain_nonconvex_vs_reference(Context, reference):
    """An L-shaped (non-convex) domain meshed with tetrahedra: per seed, the integrals over cell /\\ domain
    against the unmodified reference clipping the same cell by every tetrahedron with its own ConvexCell
    (oracle B with the tetrahedron's planes) and summing -- what generic_RVD.h:1464-1597 computes."""
    n = 6000
    xyz = oracles.uniform_seeds(n, 43)
    V, T = _lattice_mesh(2, lambda i, j, k: not (i == 1 and j == 1 and k == 1))      # the cube minus one octant
    ctx = Context(0)
    ctx.set_domain_tets(V, T)
    ctx.set_points(xyz)
    ctx.compute_cells()
    m, mg, st = ctx.get_cells()
    assert ((st & ~np.uint8(4)) == 0).all(), dict(zip(*np.unique(st, return_counts=True)))
    assert abs(m.sum() - 0.875) < 1e-10
    mr = np.zeros(n)
    mgr = np.zeros((n, 3))
    box = [-0.1, -0.12, -0.14, 1.1, 1.12, 1.14]
    for t in T:
        P = V[t]
        planes = []
        for f in range(4):
            a, b, c, o = P[(f + 1) % 4], P[(f + 2) % 4], P[(f + 3) % 4], P[f]
            nrm = np.cross(b - a, c - a)
            d = -nrm @ a
            if nrm @ o + d < 0:
                nrm, d = -nrm, -d
            planes.append([nrm[0], nrm[1], nrm[2], d])
        A = reference.cells_knn(xyz, box=box, domain=np.array(planes))
        mr += A.m
        mgr += A.mg
    assert np.abs(m - mr).max() <= 1e-9 * mr.max()
    assert np.abs(mg - mgr).max() <= 1e-9 * mr.max()
    # seeds inside the missing octant still own the part of their cell that reaches into the domain, or nothing
    inside_hole = (xyz > 0.5).all(axis=1)
    assert (m[inside_hole] < m[~inside_hole].mean()).mean() > 0.5
    ctx.close()


def test_otm_in_convex_domain(Context):
    """Semi-discrete optimal transport with a convex polytope as the source domain: the Newton solver
    equalises the Laguerre cell volumes restricted to the polytope (faces on the domain planes are
    boundary, like box faces); the reference runs OptimalTransportMap3d on a tetrahedral mesh of such a domain
    (src/examples/exploragram/compute_OTM/main.cpp:251-298)."""
    from geogram_b200 import B200VorError
    n = 20000
    dom = _convex_domain()
    xyz = oracles.uniform_seeds(3 * n, 47)
    xyz = xyz[((xyz @ dom[:, :3].T + dom[:, 3]) >= 0.01).all(axis=1)][:n]
    n = xyz.shape[0]
    ctx = Context(0)
    ctx.set_domain_planes(dom)
    ctx.set_points(xyz)
    ctx.compute_cells()
    m0, _, st0 = ctx.get_cells()
    assert (st0 == 0).all()
    vol = m0.sum()
    with pytest.raises(B200VorError):
        ctx.otm_newton(epsilon=1e-3)                      # the measure of the domain must be given
    w, res = ctx.otm_newton(epsilon=1e-3, total_mass=vol)
    assert res["converged"] == 1 and res["max_rel_error"] < 2e-3
    assert abs(res["sum_volume"] - vol) < 1e-9
    ctx.set_points(xyz, w)
    ctx.compute_cells()
    m, _, st = ctx.get_cells()
    assert (st == 0).all()
    assert np.abs(m - vol / n).max() <= 2e-3 * vol / n
    assert np.std(m0) > 20 * np.std(m)
    ctx.close()
