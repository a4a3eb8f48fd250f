"""ctypes bindings of the two CPU checkers (TEST INFRASTRUCTURE).

* ``Port``      -- oracle/_build/libvoro_oracle.so, this repo's own CPU restatement
                   (oracle/voro_oracle.cpp), always buildable.
* ``Reference`` -- oracle/_ref/libgeoref.so, the UNMODIFIED reference compiled from
                   /root/reference by oracle/Makefile (present in the build
                   container and shipped prebuilt to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(ROOT, "oracle", "_build", "libvoro_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgeoref.so")

BOX_ID = 0xFFFFFFF0
DOMAIN_ID = 0xFFFFFF00
UNCERTIFIED = 0xFFFFFFFF
UNIT_BOX = np.array([0.0, 0.0, 0.0, 1.0, 1.0, 1.0])

_cd = ctypes.POINTER(ctypes.c_double)
_cu = ctypes.POINTER(ctypes.c_uint32)


def _p(a, t):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(t)


def build_port():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


def build_reference():
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "ref"])


def have_reference():
    return os.path.exists(REF_SO)


class CellResult(object):
    """Per-cell outputs shared by every oracle (and by the GPU wrapper)."""

    def __init__(self, n, fcap, areas):
        self.n = n
        self.fcap = fcap
        self.m = np.zeros(n)
        self.mg = np.zeros((n, 3))
        self.nfaces = np.zeros(n, np.uint32)
        self.face_nbr = np.zeros((n, fcap), np.uint32)
        self.face_area = np.zeros((n, fcap)) if areas else None
        self.kstop = np.zeros(n, np.uint32)
        self.flags = np.zeros(n, np.uint32)
        self.stats = np.zeros(6)
        self.times = np.zeros(2)

    def faces(self, i):
        return self.face_nbr[i, : self.nfaces[i]]


class Port(object):
    def __init__(self):
        if not os.path.exists(PORT_SO):
            build_port()
        self.lib = ctypes.CDLL(PORT_SO)

    def set_num_threads(self, t):
        self.lib.vo_set_num_threads(int(t))

    def num_threads(self):
        return self.lib.vo_num_threads()

    def knn(self, xyz, k, queries=None):
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        q = xyz if queries is None else np.ascontiguousarray(queries, np.float64)
        nq = q.shape[0]
        nbr = np.zeros((nq, k), np.uint32)
        d2 = np.zeros((nq, k))
        r = self.lib.vo_knn_points(
            ctypes.c_uint32(n), _p(xyz, _cd), ctypes.c_uint32(nq), _p(q, _cd),
            ctypes.c_uint32(k), _p(nbr, _cu), _p(d2, _cd))
        assert r == 0
        return nbr, d2

    def cells(self, xyz, w=None, periodic=False, period=1.0, box=UNIT_BOX,
              k0=20, kcap=1000, fcap=64, areas=False, energy=False, domain=None):
        """domain: optional (k,4) planes a x + b y + c z + d >= 0 of a convex domain (box domains)"""
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        dom = np.zeros((0, 4)) if domain is None else np.ascontiguousarray(domain, np.float64)
        self.lib.vo_set_domain_planes(ctypes.c_uint32(dom.shape[0]), _p(dom, _cd) if dom.shape[0] else None)
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
        box = np.ascontiguousarray(box, np.float64)
        R = CellResult(n, fcap, areas)
        R.energy = np.zeros(n) if energy else None
        self.lib.vo_set_energy_output(_p(R.energy, _cd))
        r = self.lib.vo_cells(
            ctypes.c_uint32(n), _p(xyz, _cd), _p(w, _cd), int(periodic),
            ctypes.c_double(period), _p(box, _cd), ctypes.c_uint32(k0),
            ctypes.c_uint32(kcap), _p(R.m, _cd), _p(R.mg, _cd), _p(R.nfaces, _cu),
            _p(R.face_nbr, _cu), _p(R.face_area, _cd), ctypes.c_uint32(fcap),
            _p(R.kstop, _cu), _p(R.flags, _cu), _p(R.stats, _cd))
        self.lib.vo_set_energy_output(None)
        self.lib.vo_set_domain_planes(ctypes.c_uint32(0), None)
        assert r == 0
        return R

    def lloyd_update(self, xyz, m, mg, periodic=False, period=1.0):
        xyz = np.ascontiguousarray(xyz, np.float64)
        out = np.zeros_like(xyz)
        self.lib.vo_lloyd_update(
            ctypes.c_uint32(xyz.shape[0]), _p(xyz, _cd), _p(m, _cd),
            _p(np.ascontiguousarray(mg), _cd), int(periodic),
            ctypes.c_double(period), _p(out, _cd))
        return out


class Reference(object):
    def __init__(self, nthreads=0):
        self.lib = ctypes.CDLL(REF_SO)
        self.lib.georef_init(int(nthreads))

    def num_threads(self):
        return self.lib.georef_nthreads()

    def knn(self, xyz, k, queries=None):
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        times = np.zeros(2)
        if queries is None:
            nbr = np.zeros((n, k), np.uint32)
            d2 = np.zeros((n, k))
            r = self.lib.georef_knn(ctypes.c_uint32(n), _p(xyz, _cd), ctypes.c_uint32(k),
                                    _p(nbr, _cu), _p(d2, _cd), _p(times, _cd))
        else:
            q = np.ascontiguousarray(queries, np.float64)
            nbr = np.zeros((q.shape[0], k), np.uint32)
            d2 = np.zeros((q.shape[0], k))
            r = self.lib.georef_knn_points(
                ctypes.c_uint32(n), _p(xyz, _cd), ctypes.c_uint32(q.shape[0]), _p(q, _cd),
                ctypes.c_uint32(k), _p(nbr, _cu), _p(d2, _cd))
        assert r == 0
        self.last_times = times
        return nbr, d2

    def cells_delaunay(self, xyz, w=None, periodic=False, period=1.0, exact=True,
                       fcap=64, areas=False):
        """Oracle A.  Returns None when the diagram has empty cells."""
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
        R = CellResult(n, fcap, areas)
        r = self.lib.georef_cells_delaunay(
            ctypes.c_uint32(n), _p(xyz, _cd), _p(w, _cd), int(periodic),
            ctypes.c_double(period), int(exact), _p(R.m, _cd), _p(R.mg, _cd),
            _p(R.nfaces, _cu), _p(R.face_nbr, _cu), _p(R.face_area, _cd),
            ctypes.c_uint32(fcap), _p(R.times, _cd))
        return R if r == 0 else None

    def cvt_energy(self, xyz, w=None, periodic=False, period=1.0):
        """Per-cell CVT energy on the exact cells (RVD.cpp:846-857).  None if empty cells."""
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
        f = np.zeros(n)
        r = self.lib.georef_cvt_energy(ctypes.c_uint32(n), _p(xyz, _cd), _p(w, _cd), int(periodic),
                                       ctypes.c_double(period), _p(f, _cd))
        return f if r == 0 else None

    def delaunay_tets(self, xyz):
        """Finite Delaunay tetrahedra (n_tets, 4) of a non-periodic point set."""
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        cap = 8 * n + 64
        tets = np.zeros((cap, 4), np.uint32)
        self.lib.georef_delaunay_tets.restype = ctypes.c_uint32
        cnt = self.lib.georef_delaunay_tets(ctypes.c_uint32(n), _p(xyz, _cd), _p(tets, _cu), ctypes.c_uint32(cap))
        assert cnt <= cap
        return tets[:cnt]

    def brio_order(self, xyz, threshold=300, ratio=0.125):
        """(order, levels) of the reference's compute_BRIO_order (mesh_reorder.cpp:1287-1315)."""
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        order = np.zeros(n, np.uint32)
        levels = np.zeros(64, np.uint32)
        self.lib.georef_brio_order.restype = ctypes.c_uint32
        cnt = self.lib.georef_brio_order(ctypes.c_uint32(n), _p(xyz, _cd), ctypes.c_uint32(threshold),
                                         ctypes.c_double(ratio), _p(order, _cu), _p(levels, _cu), ctypes.c_uint32(64))
        assert cnt <= 64
        return order, levels[:cnt].copy()

    def cells_knn(self, xyz, w=None, box=UNIT_BOX, k0=20, kcap=1000, exact=True,
                  fcap=64, areas=False, domain=None):
        """Oracle B.  domain: optional (k,4) planes of a convex domain, clipped by the reference's own
        ConvexCell::clip_by_plane before any neighbour."""
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        dom = np.zeros((0, 4)) if domain is None else np.ascontiguousarray(domain, np.float64)
        self.lib.georef_set_domain_planes(ctypes.c_uint32(dom.shape[0]), _p(dom, _cd) if dom.shape[0] else None)
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
        box = np.ascontiguousarray(box, np.float64)
        R = CellResult(n, fcap, areas)
        r = self.lib.georef_cells_knn(
            ctypes.c_uint32(n), _p(xyz, _cd), _p(w, _cd), _p(box, _cd),
            ctypes.c_uint32(k0), ctypes.c_uint32(kcap), int(exact), _p(R.m, _cd),
            _p(R.mg, _cd), _p(R.nfaces, _cu), _p(R.face_nbr, _cu), _p(R.face_area, _cd),
            ctypes.c_uint32(fcap), _p(R.kstop, _cu), _p(R.stats, _cd), _p(R.times, _cd))
        self.lib.georef_set_domain_planes(ctypes.c_uint32(0), None)
        assert r == 0
        return R


def face_mismatches(A, B):
    """Number of cells whose sorted face lists differ."""
    bad = A.nfaces != B.nfaces
    fc = min(A.fcap, B.fcap)
    col = np.arange(fc)[None, :]
    mask = col < np.minimum(A.nfaces, B.nfaces)[:, None]
    diff = ((A.face_nbr[:, :fc] != B.face_nbr[:, :fc]) & mask).any(axis=1)
    return int((bad | diff).sum())


def translation_to_instance(tx, ty, tz):
    """delaunay/periodic.h:104-113 + periodic.cpp:80-84 (reorder_instances): 13 -> 0, 0..12 -> 1..13"""
    i = (int(tz) + 1) + 3 * (int(ty) + 1) + 9 * (int(tx) + 1)
    return 0 if i == 13 else (i + 1 if i < 13 else i)


def uniform_seeds(n, seed=12345):
    return np.random.default_rng(seed).random((n, 3))


def clustered_seeds(n, seed=12345, nclusters=32):
    """Mixture of isotropic Gaussians rejected to (0,1)^3."""
    rng = np.random.default_rng(seed)
    c = 0.1 + 0.8 * rng.random((nclusters, 3))
    s = 0.005 + 0.045 * rng.random(nclusters)
    out = np.zeros((0, 3))
    while out.shape[0] < n:
        k = rng.integers(0, nclusters, size=n)
        p = c[k] + s[k, None] * rng.standard_normal((n, 3))
        p = p[((p > 0) & (p < 1)).all(axis=1)]
        out = np.concatenate([out, p])
    return np.ascontiguousarray(out[:n])
