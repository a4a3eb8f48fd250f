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
