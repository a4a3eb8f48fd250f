"""Host-side cell export from the GPU buffers.

Inputs are the arrays of ``Context.get_faces()`` and ``Context.get_vertices()`` after
``compute_cells(want_faces=2)``:

* ``cell_loops``      -- the facets of one cell as oriented vertex loops (what
                         ``ConvexCell::save`` walks through ``triangle_adjacent``,
                         voronoi/convex_cell.cpp:334-346)
* ``save_cells_obj``  -- ``PeriodicDelaunay3d::save_cells`` / ``ConvexCell::save``
                         (delaunay/periodic_delaunay_3d.cpp:4363-4392, convex_cell.cpp:282-352):
                         one Wavefront OBJ with every cell shrunk towards its centroid
* ``delaunay_tets``   -- the Delaunay tetrahedra (``Delaunay::cell_vertex``, delaunay/delaunay.h:336-485):
                         one per cell vertex = the seed and the three neighbours whose faces meet there
* ``tet_adjacency``   -- ``Delaunay::cell_adjacent`` (delaunay/delaunay.h:369-380): the tetrahedron across
                         the facet opposite to each local vertex, -1 on the boundary

Pure index bookkeeping on the host; the geometry (vertex coordinates, face lists) was computed on
the device.
"""
import numpy as np

BOX_ID = 0xFFFFFFF0
DOMAIN_ID = 0xFFFFFF00
NOT_AVAILABLE = 0xFFFFFFFF


def cell_loops(nv, faces, nf):
    """Facets of one cell.  nv: number of vertices, faces: (>=nv, 3) positions in the face list of
    the three faces meeting at each vertex (same orientation for all vertices), nf: number of faces.
    Returns a list of nf loops (lists of vertex indices, consistently oriented)."""
    by_face = [[] for _ in range(nf)]
    for v in range(nv):
        a, b, c = (int(x) for x in faces[v])
        # rotations of the triple with each face first: (f, next, prev)
        by_face[a].append((v, b, c))
        by_face[b].append((v, c, a))
        by_face[c].append((v, a, b))
    loops = []
    for f in range(nf):
        ring = by_face[f]
        if len(ring) < 3:
            loops.append([v for v, _, _ in ring])
            continue
        # the vertex after (v, g, h) around f is the one that enters through h: (v', h, *)
        enter = {g: (v, h) for v, g, h in ring}
        v0, g0, h0 = ring[0]
        loop, h = [v0], h0
        while len(loop) < len(ring):
            if h not in enter:
                break
            v, h = enter[h]
            if v == v0:
                break
            loop.append(v)
        loops.append(loop)
    return loops


def _available(nverts, vcap):
    return (nverts != NOT_AVAILABLE) & (nverts > 0) & (nverts <= vcap)


def cell_volume_from_loops(xyz, loops, seed):
    """Volume of a cell from its exported geometry (fan of tetrahedra with the seed as apex):
    a check of vertex coordinates AND loop combinatorics."""
    vol = 0.0
    for loop in loops:
        if len(loop) < 3:
            continue
        p0 = xyz[loop[0]] - seed
        for k in range(1, len(loop) - 1):
            vol += np.dot(p0, np.cross(xyz[loop[k]] - seed, xyz[loop[k + 1]] - seed))
    return abs(vol) / 6.0


def save_cells_obj(path, nverts, vert_xyz, vert_faces, nfaces, m, mg, shrink=0.1, cells=None):
    """One OBJ file with every listed cell, each shrunk by `shrink` towards its centroid
    (ConvexCell::save: p <- shrink*g + (1-shrink)*p).  Returns (#vertices, #facets) written."""
    vcap = vert_xyz.shape[1]
    ok = _available(nverts, vcap)
    if cells is None:
        cells = np.nonzero(ok)[0]
    v_off, nfac = 1, 0
    with open(path, "w") as out:
        for i in cells:
            if not ok[i]:
                continue
            nv = int(nverts[i])
            g = mg[i] / m[i]
            P = shrink * g + (1.0 - shrink) * vert_xyz[i, :nv]
            for p in P:
                out.write("v %.17g %.17g %.17g\n" % (p[0], p[1], p[2]))
            for loop in cell_loops(nv, vert_faces[i], int(nfaces[i])):
                if len(loop) >= 3:
                    out.write("f " + " ".join(str(v + v_off) for v in loop) + "\n")
                    nfac += 1
            v_off += nv
    return v_off - 1, nfac


def save_cell_obj(nv, xyz, faces, nf, v_offset=1, shrink=0.0, g=None):
    """The OBJ text ConvexCell::save (voronoi/convex_cell.cpp:282-352) writes for the cell once it is
    re-hydrated into a VBW::ConvexCell the way the geogram plugin does it (dual triangle t = cell
    vertex t, created in order; local vertex 1 + lv = face lv): byte for byte --
    * triangles are pushed at the FRONT of the valid list (convex_cell.h:1143-1160), so the "v"
      records come in reverse creation order;
    * the loop of face v starts at the last listed triangle that touches v and follows
      triangle_adjacent(t, (lv + 1) % 3), where the triangle across the edge opposite to local
      vertex le owns the reversed directed edge (connect_triangles, convex_cell.cpp:1348-1411);
    * numbers go through operator<< of a default ostream (6 significant digits, %g)."""
    tri = [tuple(1 + int(x) for x in faces[t]) for t in range(nv)]
    owner = {}
    for t, (i, j, k) in enumerate(tri):
        owner[(j, k)] = t
        owner[(k, i)] = t
        owner[(i, j)] = t
    adj = [(owner[(k, j)], owner[(i, k)], owner[(j, i)]) for (i, j, k) in tri]
    out = []
    t_index = {}
    v2t = {}
    for pos, t in enumerate(range(nv - 1, -1, -1)):
        p = np.asarray(xyz[t], np.float64)
        if shrink != 0.0:
            p = shrink * np.asarray(g) + (1.0 - shrink) * p
        out.append("v %g %g %g\n" % (p[0], p[1], p[2]))
        t_index[t] = pos
        for v in tri[t]:
            v2t[v] = t
    for v in range(1, nf + 1):
        if v not in v2t:
            continue
        line = "f "
        t = v2t[v]
        while True:
            line += "%d " % (t_index[t] + v_offset)
            lv = tri[t].index(v)
            t = adj[t][(lv + 1) % 3]
            if t == v2t[v]:
                break
        out.append(line + "\n")
    return "".join(out)


def instance_translations():
    """instance -> integer translation (delaunay/periodic.cpp:44-86 with reorder_instances):
    instance 0 is the zero translation."""
    T = np.zeros((27, 3), np.int64)
    for tx in (-1, 0, 1):
        for ty in (-1, 0, 1):
            for tz in (-1, 0, 1):
                i = (tz + 1) + 3 * (ty + 1) + 9 * (tx + 1)
                T[0 if i == 13 else (i + 1 if i < 13 else i)] = (tx, ty, tz)
    return T


def _translation_to_instance(t):
    """integer translations (..., 3) in {-1,0,1}^3 -> instance numbers (periodic.h:104-113)"""
    i = (t[..., 2] + 1) + 3 * (t[..., 1] + 1) + 9 * (t[..., 0] + 1)
    return np.where(i == 13, 0, np.where(i < 13, i + 1, i))


def delaunay_tets(nverts, vert_faces, nfaces, face_nbr, n, xyz=None, period=0.0):
    """Delaunay tetrahedra as an (nt, 4) uint32 array of vertex ids (periodic: real + n*instance, as
    PeriodicDelaunay3d numbers translated copies).  A tetrahedron is reported by the cell of its
    smallest real vertex index (in the frame of that seed: the seed is instance 0); vertices on the
    clipping box belong to no tetrahedron, so on a box domain the tetrahedra whose circumcentre the
    box clipped away are missing.  Cells whose vertex list is not available (nverts == 0xFFFFFFFF:
    finished by the kNN-table pipeline, or more than vcap vertices) contribute nothing: their number
    is returned by `unavailable_cells`, and a caller that needs the complete triangulation must
    check it.  With `xyz` (and `period` on a periodic domain) the tetrahedra are positively oriented
    (orient3d > 0, the reference's convention, delaunay/delaunay.h:336-368)."""
    vcap = vert_faces.shape[1]
    ok = _available(nverts, vcap)
    idx = np.arange(vcap)[None, :]
    valid = ok[:, None] & (idx < nverts[:, None])
    rows = np.nonzero(valid)
    f = vert_faces[rows].astype(np.int64)                     # (K, 3) face positions
    nb = face_nbr[rows[0][:, None], f].astype(np.int64)       # (K, 3) neighbour ids
    interior = (nb < DOMAIN_ID).all(axis=1)    # no face on the box or on a plane of the convex domain
    seed = rows[0][interior]
    nb = nb[interior]
    keep = (seed[:, None] < (nb % n)).all(axis=1)
    tets = np.concatenate([seed[keep][:, None], nb[keep]], axis=1)
    if xyz is not None and len(tets):
        T = instance_translations().astype(np.float64) * float(period)
        P = np.asarray(xyz)[tets % n] + T[tets // n]
        det = np.einsum("ij,ij->i", P[:, 1] - P[:, 0], np.cross(P[:, 2] - P[:, 0], P[:, 3] - P[:, 0]))
        flip = det < 0
        tets[flip, 2], tets[flip, 3] = tets[flip, 3].copy(), tets[flip, 2].copy()
    return tets.astype(np.uint32)


def unavailable_cells(nverts, vcap=64):
    """Number of cells without a vertex list (they are missing from delaunay_tets / save_cells_obj):
    nverts == 0xFFFFFFFF (finished by the kNN-table pipeline) or more than vcap vertices."""
    nverts = np.asarray(nverts)
    return int(((nverts == NOT_AVAILABLE) | ((nverts != NOT_AVAILABLE) & (nverts > vcap))).sum())


def tet_adjacency(tets, n=None):
    """adj[c, lf] = tetrahedron sharing with c the facet opposite to local vertex lf, or -1
    (Delaunay::cell_adjacent, delaunay/delaunay.h:369-380).  Pure index work: facets are matched by
    their sorted vertex triples.

    Periodic domain (pass n = number of real seeds): a tetrahedron is written in the frame of the cell
    that reported it, so the same facet carries different instance numbers in its two tetrahedra;
    every facet is first brought to a canonical frame -- translated so that its smallest real vertex
    (ties: smallest translation) is instance 0 -- and then matched.  A periodic triangulation has no
    boundary: no entry is -1.  Box domain: -1 means convex hull OR a neighbour that delaunay_tets does
    not export (circumcentre clipped away by the box)."""
    tets = np.asarray(tets, np.int64)
    nt = tets.shape[0]
    adj = np.full((nt, 4), -1, np.int64)
    if nt == 0:
        return adj
    # facet lf of tet c = its three vertices other than tets[c, lf]
    lf = np.tile(np.arange(4), nt)
    c = np.repeat(np.arange(nt), 4)
    faces = np.stack([np.delete(tets, k, axis=1) for k in range(4)], axis=1).reshape(-1, 3)
    if n is not None:
        T = instance_translations()
        real, tr = faces % n, T[faces // n]                                  # (F,3), (F,3,3)
        # anchor = lexicographically smallest (real, tx, ty, tz)
        code = ((real * 3 + tr[..., 0] + 1) * 3 + tr[..., 1] + 1) * 3 + tr[..., 2] + 1
        anchor = np.argmin(code, axis=1)
        rel = tr - tr[np.arange(len(faces)), anchor][:, None, :]
        bad = (np.abs(rel) > 1).any(axis=(1, 2))                              # not representable: left unmatched
        rel = np.clip(rel, -1, 1)
        faces = real + n * _translation_to_instance(rel)
        faces[bad] = -1 - np.arange(int(bad.sum()))[:, None]                  # unique keys
    faces = np.sort(faces, axis=1)
    order = np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))
    f = faces[order]
    same = (f[1:] == f[:-1]).all(axis=1)
    # a facet belongs to at most two tetrahedra of a valid triangulation
    assert not (same[1:] & same[:-1]).any(), "a facet is shared by more than two tetrahedra"
    a, b = order[:-1][same], order[1:][same]
    adj[c[a], lf[a]] = c[b]
    adj[c[b], lf[b]] = c[a]
    return adj
