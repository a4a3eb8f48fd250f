"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libgeoref.so,
built from /root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

The fixtures pin (a) the CPU restatement oracle/voro_oracle.cpp (tests/test_oracle.py)
and (b) the CUDA path (tests/test_gpu_parity.py) to outputs of the reference itself,
because the reference ships no known-answer vectors for this path.
Face lists are stored as CSR (ptr, ids) to keep the files small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracles  # noqa: E402


def csr(R):
    ptr = np.zeros(R.n + 1, np.int64)
    ptr[1:] = np.cumsum(R.nfaces)
    ids = np.concatenate([R.faces(i) for i in range(R.n)]).astype(np.uint32)
    return ptr, ids


def save_cells(name, xyz, R, w=None, **meta):
    ptr, ids = csr(R)
    d = dict(xyz=xyz, m=R.m, mg=R.mg, face_ptr=ptr, face_ids=ids)
    if R.face_area is not None:
        d["face_area"] = np.concatenate([R.face_area[i, : R.nfaces[i]] for i in range(R.n)])
    if w is not None:
        d["w"] = w
    for k, v in meta.items():
        d[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **d)
    print(name, "n=%d sum(m)=%.15f mean faces=%.3f" % (R.n, R.m.sum(), R.nfaces.mean()))


def main():
    oracles.build_reference()
    R = oracles.Reference()
    # kNN: BalancedKdTree, by index (k=21 = reference default 20 + self) and by point
    xyz = oracles.uniform_seeds(3000, 12345)
    nbr, d2 = R.knn(xyz, 21)
    q = np.random.default_rng(7).random((500, 3)) * 1.2 - 0.1
    nbq, d2q = R.knn(xyz, 9, queries=q)
    np.savez_compressed(os.path.join(HERE, "knn_uniform_3000.npz"), xyz=xyz, nbr=nbr, d2=d2,
                        q=q, nbr_q=nbq, d2_q=d2q)
    print("knn_uniform_3000.npz")
    # config 1 shape: uniform seeds, unit box.  Oracle A (exact Delaunay) and oracle B
    # (kNN + VBW::ConvexCell + security radius, k0=20) must agree; store A + B's stats.
    A = R.cells_delaunay(xyz, areas=True)
    B = R.cells_knn(xyz, k0=20)
    assert oracles.face_mismatches(A, B) == 0
    save_cells("cells_box_uniform_3000.npz", xyz, A, kstop_B=B.kstop, stats_B=B.stats)
    # config 2 shape: periodic unit cube (ids = real + n*instance)
    A = R.cells_delaunay(xyz, periodic=True, areas=True)
    save_cells("cells_periodic_uniform_3000.npz", xyz, A)
    # config 3 shape: Laguerre cells, periodic, w = a h^2 U(0,1), a = 0.02 (non-emptying)
    n = xyz.shape[0]
    h = n ** (-1.0 / 3.0)
    w = 0.02 * h * h * np.random.default_rng(5).random(n)
    A = R.cells_delaunay(xyz, w=w, periodic=True, areas=True)
    assert A is not None
    save_cells("cells_periodic_laguerre_3000.npz", xyz, A, w=w)
    # Laguerre in a box, larger weights on Lloyd-relaxed seeds is covered on the GPU side;
    # here: box + weights
    A = R.cells_delaunay(xyz, w=w, areas=True)
    assert A is not None
    save_cells("cells_box_laguerre_3000.npz", xyz, A, w=w)
    # config 4 shape: clustered seeds in a box (kNN + SR alone cannot certify the fringe)
    xc = oracles.clustered_seeds(4000, 12345)
    A = R.cells_delaunay(xc, areas=False)
    save_cells("cells_box_clustered_4000.npz", xc, A)
    # two Lloyd iterations, periodic (demo_Delaunay3d/main.cpp:432-456 update rule)
    x = xyz.copy()
    P = oracles.Port()
    for _ in range(2):
        A = R.cells_delaunay(x, periodic=True)
        x = P.lloyd_update(x, A.m, A.mg, True, 1.0)
    np.savez_compressed(os.path.join(HERE, "lloyd2_periodic_3000.npz"), xyz0=xyz, xyz2=x)
    print("lloyd2_periodic_3000.npz")


def cvt_energy():
    """CVT objective (ComputeCVTFuncGradVolumetric, RVD.cpp:791-876) on the exact cells of the
    unmodified reference: f_i per cell (georef_cvt_energy) and g_i = 2 (m_i p_i - mg_i) from
    the reference's own compute_mg."""
    R = oracles.Reference()
    xyz = oracles.uniform_seeds(3000, 12345)
    for periodic, tag in ((False, "box"), (True, "periodic")):
        A = R.cells_delaunay(xyz, periodic=periodic)
        f = R.cvt_energy(xyz, periodic=periodic)
        g = 2.0 * (A.m[:, None] * xyz - A.mg)
        np.savez_compressed(os.path.join(HERE, "cvt_energy_%s_3000.npz" % tag), xyz=xyz, f_cell=f, g=g)
        print("cvt_energy_%s_3000.npz f=%.15e" % (tag, f.sum()))


def delaunay_tets():
    """Finite Delaunay tetrahedra of the unmodified reference (Delaunay::cell_vertex) on the
    3000-seed box set: the pin of the tetrahedra exported from the GPU cell vertices."""
    R = oracles.Reference()
    xyz = oracles.uniform_seeds(3000, 12345)
    t = np.sort(R.delaunay_tets(xyz), axis=1)
    np.savez_compressed(os.path.join(HERE, "delaunay_tets_box_3000.npz"), xyz=xyz, tets=t)
    print("delaunay_tets_box_3000.npz", t.shape)


def lattice_points(k=12, cells=16):
    """k^3 lattice with DYADIC coordinates (i + 1/2) / cells: every bisector coefficient is exact in
    fp64 in any frame, so the exact-mode decisions of the reference and of the GPU path are decisions
    about the same real numbers."""
    g = (np.arange(k) + 0.5) / float(cells)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    return np.ascontiguousarray(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1))


def lattice():
    """Degenerate input: a cubic lattice in the box [0, 0.75]^3, cells by the unmodified reference's
    kNN + VBW::ConvexCell path in its DEFAULT exact mode (ties are conflicts, convex_cell.cpp:905-911):
    26 faces per interior cell, 18 of them of zero area."""
    R = oracles.Reference()
    xyz = lattice_points()
    box = np.array([0, 0, 0, 0.75, 0.75, 0.75])
    B = R.cells_knn(xyz, box=box, exact=True, areas=True)
    save_cells("cells_box_lattice_exact_1728.npz", xyz, B, box=box)


def brio():
    """BRIO levels of the unmodified reference (compute_BRIO_order, mesh_reorder.cpp:1287-1315) for several
    sizes / thresholds / ratios, and its order on 20 000 seeds (for the statistics of the prefixes and the
    locality of the last interval, which geogram_b200/ordering.py must match in kind)."""
    R = oracles.Reference()
    cases, out = [(20000, 300, 0.125), (4000, 300, 0.125), (301, 300, 0.125), (299, 300, 0.125),
                  (123457, 64, 0.25), (1000000, 300, 0.125)], {}
    for n, th, r in cases:
        xyz = oracles.uniform_seeds(n, 777)
        order, levels = R.brio_order(xyz, th, r)
        assert np.array_equal(np.sort(order), np.arange(n))
        out["levels_%d_%d_%g" % (n, th, r)] = levels
        if n == 20000:
            out["xyz_20000"] = xyz
            out["order_20000"] = order
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "brio_reference.npz"), **out)
    print("brio_reference.npz", {k: v.tolist() for k, v in out.items() if k.startswith("levels")})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "brio":
        brio()
    elif len(sys.argv) > 1 and sys.argv[1] == "lattice":
        lattice()
    elif len(sys.argv) > 1 and sys.argv[1] == "cvt":
        cvt_energy()
    elif len(sys.argv) > 1 and sys.argv[1] == "tets":
        delaunay_tets()
    else:
        main()
        cvt_energy()
        delaunay_tets()
        lattice()
