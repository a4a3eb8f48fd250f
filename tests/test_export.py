"""Cell export from the GPU buffers: facet loops, OBJ, Delaunay tetrahedra."""
import os

import numpy as np
import pytest

import oracles
from conftest import Golden
from geogram_b200 import export


def _cube():
    """The clipping box as the fused kernel initialises it (init_with_box, convex_cell.cpp:175-215)"""
    faces, xyz = [], []
    for s in range(8):
        bx, by, bz = s & 1, (s >> 1) & 1, (s >> 2) & 1
        px, py, pz = bx, 2 + by, 4 + bz
        v1, v2 = (pz, py) if (bx + by + bz) & 1 else (py, pz)
        faces.append((px, v1, v2))
        xyz.append((bx, by, bz))
    return np.array(faces, np.uint8), np.array(xyz, float)


def test_cell_loops_on_the_box():
    faces, xyz = _cube()
    loops = export.cell_loops(8, faces, 6)
    assert sorted(len(l) for l in loops) == [4] * 6
    assert sorted(v for l in loops for v in l) == sorted(list(range(8)) * 3)
    assert abs(export.cell_volume_from_loops(xyz, loops, np.array([0.3, 0.4, 0.5])) - 1.0) < 1e-15


def test_delaunay_tets_dedup_rule():
    # two cells that see the same tetrahedron (0,1,2,3): only cell 0 reports it
    nverts = np.array([1, 1, 0xFFFFFFFF, 0], np.uint32)
    vert_faces = np.zeros((4, 64, 3), np.uint8)
    vert_faces[0, 0] = vert_faces[1, 0] = (0, 1, 2)
    nfaces = np.array([3, 3, 0, 0], np.uint32)
    face_nbr = np.zeros((4, 64), np.uint32)
    face_nbr[0, :3] = (1, 2, 3)
    face_nbr[1, :3] = (0, 2, 3)
    t = export.delaunay_tets(nverts, vert_faces, nfaces, face_nbr, 4)
    assert t.tolist() == [[0, 1, 2, 3]]
    face_nbr[0, 2] = export.BOX_ID + 1          # a vertex on the clipping box is no tetrahedron
    assert len(export.delaunay_tets(nverts, vert_faces, nfaces, face_nbr, 4)) == 0


def test_tet_adjacency_on_reference_tets():
    """Delaunay::cell_adjacent rebuilt from the tetrahedra: on the Delaunay triangulation of the
    unmodified reference (golden) adjacency is symmetric, adjacent tetrahedra share exactly three
    vertices, and the boundary facets are those of the convex hull (Euler: interior facets =
    (4 nt - hull) / 2)."""
    d = np.load(os.path.join(oracles.ROOT, "tests", "golden", "delaunay_tets_box_3000.npz"))
    tets = d["tets"].astype(np.int64)
    adj = export.tet_adjacency(tets)
    nt = len(tets)
    c, lf = np.nonzero(adj >= 0)
    nb = adj[c, lf]
    assert ((adj[nb] == c[:, None]).sum(axis=1) == 1).all()                  # symmetric
    shared = np.array([len(set(tets[a]) & set(tets[b])) for a, b in zip(c[:500], nb[:500])])
    assert (shared == 3).all()
    hull = int((adj < 0).sum())
    assert hull > 0 and (4 * nt - hull) % 2 == 0
    # every hull facet lies on the convex hull: all seeds on one side of its plane
    xyz = d["xyz"]
    cc, ll = np.nonzero(adj < 0)
    for k in range(0, len(cc), max(1, len(cc) // 40)):
        f = np.delete(tets[cc[k]], ll[k])
        n = np.cross(xyz[f[1]] - xyz[f[0]], xyz[f[2]] - xyz[f[0]])
        s = (xyz - xyz[f[0]]) @ n
        assert (s <= 1e-12).all() or (s >= -1e-12).all()


@pytest.mark.gpu
@pytest.mark.parametrize("periodic", [True, False])
def test_export_geometry_on_gpu(tmp_path, periodic):
    from geogram_b200 import Context
    G = Golden("cells_periodic_uniform_3000.npz" if periodic else "cells_box_uniform_3000.npz")
    n = G.n
    ctx = Context(0)
    if periodic:
        ctx.set_domain_periodic(1.0)
    ctx.set_points(G.xyz)
    ctx.compute_cells(want_faces=2)
    m, mg, st = ctx.get_cells()
    nf, fn, fa = ctx.get_faces()
    nv, vx, vf = ctx.get_vertices()
    assert (st == 0).all()
    assert (nv == 2 * nf - 4).all()                      # Euler, every vertex of degree 3
    T = oracles_translations()
    worst_vol, worst_eq = 0.0, 0.0
    for i in range(0, n, 7):
        loops = export.cell_loops(int(nv[i]), vf[i], int(nf[i]))
        assert sum(len(l) for l in loops) == 3 * nv[i]
        vol = export.cell_volume_from_loops(vx[i, : nv[i]], loops, G.xyz[i])
        worst_vol = max(worst_vol, abs(vol - m[i]) / m[i])
        # a cell vertex is equidistant from the seed and the three neighbours whose faces meet there
        for v in range(int(nv[i])):
            ids = fn[i, vf[i, v]].astype(np.int64)
            p = vx[i, v]
            d0 = ((p - G.xyz[i]) ** 2).sum()
            for j in ids[ids < export.BOX_ID]:
                q = G.xyz[j % n] + T[j // n]
                worst_eq = max(worst_eq, abs(((p - q) ** 2).sum() - d0) / d0)
    assert worst_vol < 1e-9 and worst_eq < 1e-9
    # Delaunay tetrahedra
    tets = export.delaunay_tets(nv, vf, nf, fn, n)
    if periodic:
        assert 4 * len(tets) == int(nv.sum())            # every tetrahedron is seen from its 4 cells
    assert len(np.unique(np.sort(tets, axis=1), axis=0)) == len(tets)
    rng = np.random.default_rng(0)
    allp = (G.xyz[None, :, :] + T[:, None, :]).reshape(-1, 3) if periodic else G.xyz
    for t in tets[rng.integers(0, len(tets), 60)]:
        P = G.xyz[t % n] + T[t // n]
        A = 2.0 * (P[1:] - P[0])
        b = (P[1:] ** 2).sum(axis=1) - (P[0] ** 2).sum()
        c = np.linalg.solve(A, b)
        r2 = ((c - P[0]) ** 2).sum()
        d2 = ((allp - c) ** 2).sum(axis=1)
        assert (d2 >= r2 * (1 - 1e-9)).all()             # empty circumsphere
    # Delaunay::cell_adjacent on the EXPORTED tetrahedra; positive orientation
    assert export.unavailable_cells(nv) == 0
    tets = export.delaunay_tets(nv, vf, nf, fn, n, xyz=G.xyz, period=1.0 if periodic else 0.0)
    P = G.xyz[tets % n] + T[tets // n]
    det = np.einsum("ij,ij->i", P[:, 1] - P[:, 0], np.cross(P[:, 2] - P[:, 0], P[:, 3] - P[:, 0]))
    assert (det > 0).all()
    adj = export.tet_adjacency(tets, n if periodic else None)
    c, lf = np.nonzero(adj >= 0)
    assert ((adj[adj[c, lf]] == c[:, None]).sum(axis=1) == 1).all()           # symmetric
    if periodic:
        assert (adj >= 0).all()          # a periodic triangulation has no boundary
    else:
        assert (adj >= 0).mean() > 0.7   # -1 = hull, or a neighbour clipped away by the box
    # OBJ (ConvexCell::save format)
    path = os.path.join(str(tmp_path), "cells.obj")
    nvw, nfw = export.save_cells_obj(path, nv, vx, vf, nf, m, mg, shrink=0.1, cells=np.arange(0, n, 50))
    lines = open(path).read().splitlines()
    assert sum(l.startswith("v ") for l in lines) == nvw == int(nv[0:n:50].sum())
    assert sum(l.startswith("f ") for l in lines) == nfw == int(nf[0:n:50].sum())
    ctx.close()


@pytest.mark.gpu
def test_delaunay_tets_vs_reference_golden():
    """Tetrahedra rebuilt from the GPU cell vertices against Delaunay::cell_vertex of the unmodified
    reference (golden): every exported tetrahedron is a Delaunay tetrahedron, and every Delaunay
    tetrahedron whose circumcentre lies inside the clipping box is exported (the others have their
    Voronoi vertex clipped away by the box)."""
    from geogram_b200 import Context
    d = np.load(os.path.join(oracles.ROOT, "tests", "golden", "delaunay_tets_box_3000.npz"))
    xyz, ref = d["xyz"], d["tets"]
    n = xyz.shape[0]
    ctx = Context(0)
    ctx.set_points(xyz)
    ctx.compute_cells(want_faces=2)
    nf, fn, fa = ctx.get_faces()
    nv, vx, vf = ctx.get_vertices()
    ours = np.sort(export.delaunay_tets(nv, vf, nf, fn, n), axis=1)
    key = lambda t: set(map(tuple, t.tolist()))
    S_ours, S_ref = key(ours), key(ref)
    assert S_ours <= S_ref
    P = xyz[ref]                                            # (nt, 4, 3)
    A = 2.0 * (P[:, 1:] - P[:, :1])
    b = (P[:, 1:] ** 2).sum(axis=2) - (P[:, :1] ** 2).sum(axis=2)
    cc = np.linalg.solve(A, b[..., None])[..., 0]
    inside = ((cc > 1e-9) & (cc < 1 - 1e-9)).all(axis=1)
    assert key(ref[inside]) <= S_ours
    assert len(S_ours) > 0.8 * len(S_ref)
    ctx.close()


def oracles_translations():
    """instance -> translation, delaunay/periodic.cpp:44-86 with reorder_instances"""
    T = np.zeros((27, 3))
    for tx in (-1, 0, 1):
        for ty in (-1, 0, 1):
            for tz in (-1, 0, 1):
                i = (tz + 1) + 3 * (ty + 1) + 9 * (tx + 1)
                T[0 if i == 13 else (i + 1 if i < 13 else i)] = (tx, ty, tz)
    return T


@pytest.mark.gpu
def test_hilbert_order_is_a_local_permutation():
    """b200vor_get_hilbert_order (role of compute_Hilbert_order, mesh/mesh_reorder.cpp:1260-1300): a
    permutation of the seeds in which consecutive entries are spatial neighbours EVERYWHERE -- the raster
    order of the grid jumps across the domain at the end of every row of bins, the curve never does."""
    import oracles
    from geogram_b200 import Context
    n = 200000
    xyz = oracles.uniform_seeds(n, 17)
    ctx = Context(0)
    ctx.set_points(xyz)
    h = ctx.hilbert_order()
    r = ctx.spatial_order()
    assert sorted(h.tolist()) == list(range(n))
    w = 1.0 / np.array(ctx.stats()["grid_dim"], float).min()       # bin width
    jh = np.linalg.norm(np.diff(xyz[h], axis=0), axis=1)
    jr = np.linalg.norm(np.diff(xyz[r], axis=0), axis=1)
    assert jh.max() <= 4.0 * np.sqrt(3.0) * w                      # consecutive bins are at most a bin apart (+ a rare empty bin)
    assert jr.max() > 0.5                                          # the raster order jumps across the box
    assert jh.mean() <= jr.mean()
    # small and elongated sets still give a permutation
    for pts in (oracles.uniform_seeds(7, 1), oracles.uniform_seeds(3000, 2) * np.array([1.0, 0.01, 0.01])):
        ctx.set_points(pts)
        assert sorted(ctx.hilbert_order().tolist()) == list(range(pts.shape[0]))
    ctx.close()
