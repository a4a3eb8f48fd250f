"""ctypes binding of include/b200vor.h (libb200vor.so).

Thin by design: every method is one C-ABI call.  There is no CPU fallback: if the
shared library is missing, or no CUDA device is present, the call raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200VOR_LIB") or os.path.join(_HERE, "lib", "libb200vor.so")

BOX_ID = 0xFFFFFFF0
DOMAIN_ID = 0xFFFFFF00
CELL_OK = 0
CELL_UNCERTIFIED = 1
CELL_UNCERTAIN = 2
CELL_EMPTY = 4
CELL_BOXFACE = 8
CELL_OVERFLOW = 16
CELL_DUPLICATE = 32
CELL_DEGENERATE = 64
FCAP = 64

# every symbol include/b200vor.h declares
EXPORTS = [
    "b200vor_version", "b200vor_last_error", "b200vor_create", "b200vor_destroy",
    "b200vor_set_domain_box", "b200vor_set_domain_periodic", "b200vor_set_domain_planes", "b200vor_set_domain_pieces", "b200vor_set_domain_tets", "b200vor_set_points",
    "b200vor_knn", "b200vor_knn_points", "b200vor_compute_cells", "b200vor_get_cells",
    "b200vor_get_faces", "b200vor_get_vertices", "b200vor_get_cvt", "b200vor_otm_newton", "b200vor_get_spatial_order", "b200vor_get_hilbert_order", "b200vor_lloyd", "b200vor_get_stats",
    "b200vor_set_points_per_cell", "b200vor_set_stream", "b200vor_set_query_range",
    "b200vor_microbench", "b200vor_selftest_exact_sign",
    "b200vor_comm_unique_id", "b200vor_comm_init", "b200vor_slab_set_halo_bins", "b200vor_slab_set_points",
    "b200vor_slab_compute_cells", "b200vor_slab_lloyd", "b200vor_slab_get_stats",
]
UNIQUE_ID_BYTES = 128


class CellParams(ctypes.Structure):
    _fields_ = [("k0", ctypes.c_uint32), ("kmax", ctypes.c_uint32),
                ("want_faces", ctypes.c_uint32), ("mode", ctypes.c_uint32),
                ("want_energy", ctypes.c_uint32)]


class OtmParams(ctypes.Structure):
    _fields_ = [("max_newton", ctypes.c_uint32), ("max_cg", ctypes.c_uint32),
                ("cg_tol", ctypes.c_double), ("epsilon", ctypes.c_double),
                ("total_mass", ctypes.c_double)]


class OtmResult(ctypes.Structure):
    _fields_ = [("newton_iters", ctypes.c_uint32), ("cg_iters", ctypes.c_uint32),
                ("cell_evals", ctypes.c_uint32), ("converged", ctypes.c_int32),
                ("initial_error", ctypes.c_double), ("final_error", ctypes.c_double),
                ("max_rel_error", ctypes.c_double), ("min_volume", ctypes.c_double),
                ("sum_volume", ctypes.c_double), ("last_alpha", ctypes.c_double)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class Stats(ctypes.Structure):
    _fields_ = [
        ("ms_grid", ctypes.c_float), ("ms_knn", ctypes.c_float),
        ("ms_clip", ctypes.c_float), ("ms_total", ctypes.c_float),
        ("passes", ctypes.c_uint32),
        ("n_pass", ctypes.c_uint32 * 8), ("k_pass", ctypes.c_uint32 * 8),
        ("n_uncertified", ctypes.c_uint32), ("n_uncertain", ctypes.c_uint32),
        ("n_empty", ctypes.c_uint32), ("n_boxface", ctypes.c_uint32),
        ("n_overflow", ctypes.c_uint32), ("n_duplicate", ctypes.c_uint32),
        ("n_degenerate", ctypes.c_uint32),
        ("sum_planes", ctypes.c_uint64), ("sum_tests", ctypes.c_uint64),
        ("sum_created", ctypes.c_uint64), ("sum_tets", ctypes.c_uint64),
        ("gpu_launches", ctypes.c_uint64),
        ("grid_dim", ctypes.c_uint32 * 3),
        ("tpc_cells", ctypes.c_uint32),
        ("sum_scans", ctypes.c_uint64),
        ("exact_cells", ctypes.c_uint32), ("reserved1", ctypes.c_uint32),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


class SlabStats(ctypes.Structure):
    _fields_ = [("n_own", ctypes.c_uint32), ("n_ghost_x", ctypes.c_uint32), ("n_ghost_yz", ctypes.c_uint32),
                ("n_local", ctypes.c_uint32), ("retries", ctypes.c_uint32), ("n_flagged", ctypes.c_uint32),
                ("halo", ctypes.c_double), ("slab_lo", ctypes.c_double), ("slab_hi", ctypes.c_double),
                ("ms_exchange", ctypes.c_float), ("ms_pack", ctypes.c_float), ("ms_grid", ctypes.c_float),
                ("ms_clip", ctypes.c_float), ("bytes_sent", ctypes.c_uint64), ("bytes_received", ctypes.c_uint64),
                ("reallocs", ctypes.c_uint64)]


class Peaks(ctypes.Structure):
    _fields_ = [("fp64_tflops", ctypes.c_double), ("fp32_tflops", ctypes.c_double),
                ("issue_ginst_s", ctypes.c_double), ("sm_clock_mhz", ctypes.c_double),
                ("sm_count", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class B200VorError(RuntimeError):
    pass


_lib = None


def load_library():
    """Loads libb200vor.so (raises if it was not built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200VorError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C geogram_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, u32, dbl, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_double, ctypes.c_int
    lib.b200vor_last_error.restype = ctypes.c_char_p
    lib.b200vor_create.argtypes = [ctypes.POINTER(vp), i32]
    lib.b200vor_destroy.argtypes = [vp]
    lib.b200vor_destroy.restype = None
    lib.b200vor_set_domain_box.argtypes = [vp, vp]
    lib.b200vor_set_domain_periodic.argtypes = [vp, dbl]
    lib.b200vor_set_domain_planes.argtypes = [vp, u32, vp]
    lib.b200vor_set_domain_pieces.argtypes = [vp, u32, vp, vp]
    lib.b200vor_set_domain_tets.argtypes = [vp, u32, vp, u32, vp]
    lib.b200vor_set_points.argtypes = [vp, u32, vp, u32, vp, i32]
    lib.b200vor_knn.argtypes = [vp, u32, vp, vp, i32]
    lib.b200vor_knn_points.argtypes = [vp, u32, vp, u32, vp, vp, i32]
    lib.b200vor_compute_cells.argtypes = [vp, ctypes.POINTER(CellParams)]
    lib.b200vor_get_cells.argtypes = [vp, vp, vp, vp, i32]
    lib.b200vor_get_faces.argtypes = [vp, vp, vp, vp, u32, i32]
    lib.b200vor_get_vertices.argtypes = [vp, vp, vp, vp, u32, i32]
    lib.b200vor_get_cvt.argtypes = [vp, vp, vp, ctypes.POINTER(dbl), i32]
    lib.b200vor_otm_newton.argtypes = [vp, vp, vp, ctypes.POINTER(OtmParams), ctypes.POINTER(OtmResult), i32]
    lib.b200vor_get_spatial_order.argtypes = [vp, vp, i32]
    lib.b200vor_get_hilbert_order.argtypes = [vp, vp, i32]
    lib.b200vor_lloyd.argtypes = [vp, u32, ctypes.POINTER(CellParams), vp, i32]
    lib.b200vor_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    lib.b200vor_set_points_per_cell.argtypes = [vp, dbl]
    lib.b200vor_set_stream.argtypes = [vp, vp]
    lib.b200vor_set_query_range.argtypes = [vp, u32, u32]
    lib.b200vor_microbench.argtypes = [vp, ctypes.POINTER(Peaks)]
    lib.b200vor_selftest_exact_sign.argtypes = [vp, u32, vp, vp]
    lib.b200vor_comm_unique_id.argtypes = [vp]
    lib.b200vor_comm_init.argtypes = [vp, vp, i32, i32]
    lib.b200vor_slab_set_halo_bins.argtypes = [vp, dbl]
    lib.b200vor_slab_set_points.argtypes = [vp, u32, ctypes.c_uint64, ctypes.c_uint64, vp, vp, i32]
    lib.b200vor_slab_compute_cells.argtypes = [vp, ctypes.POINTER(CellParams)]
    lib.b200vor_slab_lloyd.argtypes = [vp, u32, ctypes.POINTER(CellParams), vp, i32]
    lib.b200vor_slab_get_stats.argtypes = [vp, ctypes.POINTER(SlabStats)]
    _lib = lib
    return lib


def _ptr(a):
    """Host numpy array / raw device address (int) / None -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


class Context(object):
    """One engine instance on one GPU (owns its stream and device buffers)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        self._check(self.lib.b200vor_create(ctypes.byref(h), int(device)))
        self.h = h
        self.n = 0

    def _check(self, rc):
        if rc != 0:
            raise B200VorError("b200vor error %d: %s" % (
                rc, self.lib.b200vor_last_error().decode("utf-8", "replace")))

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200vor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- domain / points ----
    def set_domain_box(self, box):
        box = np.ascontiguousarray(box, np.float64)
        self._check(self.lib.b200vor_set_domain_box(self.h, _ptr(box)))

    def set_domain_planes(self, planes):
        """planes: (k,4) rows a, b, c, d of the half-spaces a x + b y + c z + d >= 0 of a convex domain
        (k <= 32; None or empty removes them); box domains only"""
        pl = np.zeros((0, 4)) if planes is None else np.ascontiguousarray(planes, np.float64).reshape(-1, 4)
        self._check(self.lib.b200vor_set_domain_planes(self.h, pl.shape[0], _ptr(pl) if pl.shape[0] else None))

    def set_domain_tets(self, vertices, tets):
        """domain = a tetrahedral mesh: vertices (nv,3) float64, tets (nt,4) indices; None removes it"""
        if vertices is None:
            self._check(self.lib.b200vor_set_domain_pieces(self.h, 0, None, None))
            return
        v = np.ascontiguousarray(vertices, np.float64)
        t = np.ascontiguousarray(tets, np.uint32)
        self._check(self.lib.b200vor_set_domain_tets(self.h, v.shape[0], _ptr(v), t.shape[0], _ptr(t)))

    def set_domain_periodic(self, period=1.0):
        self._check(self.lib.b200vor_set_domain_periodic(self.h, float(period)))

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)"""
        self._check(self.lib.b200vor_set_stream(self.h, ctypes.c_void_p(int(cuda_stream))))

    def set_query_range(self, first, last):
        self._check(self.lib.b200vor_set_query_range(self.h, int(first), int(last)))

    def set_points_per_cell(self, ppc):
        self._check(self.lib.b200vor_set_points_per_cell(self.h, float(ppc)))

    def set_points(self, xyz, w=None):
        """Host numpy arrays (copied H2D inside the call)."""
        xyz = np.ascontiguousarray(xyz, np.float64)
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
        self.n = xyz.shape[0]
        self._check(self.lib.b200vor_set_points(self.h, self.n, _ptr(xyz), 3, _ptr(w), 0))

    def set_points_device(self, n, xyz_ptr, w_ptr=None, stride=3):
        """Raw device addresses (e.g. torch.Tensor.data_ptr())."""
        self.n = int(n)
        self._check(self.lib.b200vor_set_points(
            self.h, self.n, _ptr(int(xyz_ptr)), int(stride),
            _ptr(int(w_ptr)) if w_ptr else None, 1))

    # ---- kNN ----
    def knn(self, k, want_d2=True):
        nbr = np.zeros((self.n, k), np.uint32)
        d2 = np.zeros((self.n, k)) if want_d2 else None
        self._check(self.lib.b200vor_knn(self.h, int(k), _ptr(nbr), _ptr(d2), 0))
        return nbr, d2

    def knn_device(self, k, nbr_ptr, d2_ptr=None):
        """rows of k into device buffers (n*k uint32, optional n*k float64), raw device addresses"""
        self._check(self.lib.b200vor_knn(self.h, int(k), _ptr(int(nbr_ptr)),
                                         _ptr(int(d2_ptr)) if d2_ptr else None, 1))

    def knn_points(self, q, k):
        q = np.ascontiguousarray(q, np.float64)
        nq = q.shape[0]
        nbr = np.zeros((nq, k), np.uint32)
        d2 = np.zeros((nq, k))
        self._check(self.lib.b200vor_knn_points(self.h, nq, _ptr(q), int(k), _ptr(nbr), _ptr(d2), 0))
        return nbr, d2

    # ---- cells ----
    def compute_cells(self, k0=0, kmax=0, want_faces=False, mode=0, want_energy=False):
        # want_faces: False / True (face lists) / 2 (face lists + cell vertices)
        p = CellParams(int(k0), int(kmax), int(want_faces), int(mode), 1 if want_energy else 0)
        self._check(self.lib.b200vor_compute_cells(self.h, ctypes.byref(p)))

    def get_cells(self):
        m = np.zeros(self.n)
        mg = np.zeros((self.n, 3))
        status = np.zeros(self.n, np.uint8)
        self._check(self.lib.b200vor_get_cells(self.h, _ptr(m), _ptr(mg), _ptr(status), 0))
        return m, mg, status

    def get_cells_device(self, m_ptr, mg_ptr, status_ptr):
        self._check(self.lib.b200vor_get_cells(
            self.h, _ptr(int(m_ptr)) if m_ptr else None, _ptr(int(mg_ptr)) if mg_ptr else None,
            _ptr(int(status_ptr)) if status_ptr else None, 1))

    def get_faces(self):
        nf = np.zeros(self.n, np.uint32)
        nbr = np.zeros((self.n, FCAP), np.uint32)
        area = np.zeros((self.n, FCAP))
        self._check(self.lib.b200vor_get_faces(self.h, _ptr(nf), _ptr(nbr), _ptr(area), FCAP, 0))
        return nf, nbr, area

    def get_vertices(self):
        """(nverts[n], xyz[n,64,3], faces[n,64,3]) of compute_cells(want_faces=2): cell vertices and,
        per vertex, the positions in the face list of the three faces that meet there."""
        nv = np.zeros(self.n, np.uint32)
        xyz = np.zeros((self.n, FCAP, 3))
        vf = np.zeros((self.n, FCAP), np.uint32)
        self._check(self.lib.b200vor_get_vertices(self.h, _ptr(nv), _ptr(xyz), _ptr(vf), FCAP, 0))
        faces = np.stack([vf & 0xFF, (vf >> 8) & 0xFF, (vf >> 16) & 0xFF], axis=-1).astype(np.uint8)
        return nv, xyz, faces

    def spatial_order(self):
        """permutation of the seeds in grid order (spatial sort; role of compute_Hilbert_order)"""
        order = np.zeros(self.n, np.uint32)
        self._check(self.lib.b200vor_get_spatial_order(self.h, _ptr(order), 0))
        return order

    def hilbert_order(self):
        """permutation of the seeds along the Hilbert curve of the grid (bin by bin, inside a bin by index)"""
        order = np.zeros(self.n, np.uint32)
        self._check(self.lib.b200vor_get_hilbert_order(self.h, _ptr(order), 0))
        return order

    def get_cvt(self):
        """(f_total, f_cell[n], g[n,3]) of the CVT objective (RVD.cpp:791-876); needs
        compute_cells(want_energy=True)"""
        f = np.zeros(self.n)
        g = np.zeros((self.n, 3))
        tot = ctypes.c_double(0.0)
        self._check(self.lib.b200vor_get_cvt(self.h, _ptr(f), _ptr(g), ctypes.byref(tot), 0))
        return tot.value, f, g

    def otm_newton(self, nu=None, w0=None, max_newton=0, max_cg=0, cg_tol=0.0, epsilon=0.0, total_mass=0.0):
        """Semi-discrete optimal transport (damped Newton on the Laguerre weights) for the points
        of set_points: returns (w, result dict).  nu: target measures or None (uniform)."""
        w = np.zeros(self.n) if w0 is None else np.array(w0, np.float64, copy=True)
        if nu is not None:
            nu = np.ascontiguousarray(nu, np.float64)
        p = OtmParams(int(max_newton), int(max_cg), float(cg_tol), float(epsilon), float(total_mass))
        r = OtmResult()
        self._check(self.lib.b200vor_otm_newton(self.h, _ptr(nu), _ptr(w), ctypes.byref(p), ctypes.byref(r), 0))
        return w, r.as_dict()

    def lloyd(self, iters, k0=0, kmax=0, mode=0, want_xyz=True, out=None):
        """out: optional preallocated (n,3) float64 host array (e.g. pinned) for the result"""
        p = CellParams(int(k0), int(kmax), 0, int(mode), 0)
        if out is None:
            out = np.zeros((self.n, 3)) if want_xyz else None
        self._check(self.lib.b200vor_lloyd(self.h, int(iters), ctypes.byref(p), _ptr(out), 0))
        return out

    def lloyd_device(self, iters, xyz_out_ptr=None, k0=0, kmax=0, mode=0):
        p = CellParams(int(k0), int(kmax), 0, int(mode), 0)
        self._check(self.lib.b200vor_lloyd(
            self.h, int(iters), ctypes.byref(p),
            _ptr(int(xyz_out_ptr)) if xyz_out_ptr else None, 1))

    # ---- multi-GPU slabs (include/b200vor.h, "x-slabs with halos"); one process per GPU ----
    @staticmethod
    def comm_unique_id():
        """128 bytes (an ncclUniqueId) created by ONE rank; hand them to every rank"""
        lib = load_library()
        buf = ctypes.create_string_buffer(UNIQUE_ID_BYTES)
        if lib.b200vor_comm_unique_id(buf) != 0:
            raise B200VorError(lib.b200vor_last_error().decode("utf-8", "replace"))
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        buf = ctypes.create_string_buffer(bytes(unique_id), UNIQUE_ID_BYTES) if unique_id is not None else None
        self._check(self.lib.b200vor_comm_init(self.h, buf, int(rank), int(world)))

    def slab_set_halo_bins(self, bins):
        self._check(self.lib.b200vor_slab_set_halo_bins(self.h, float(bins)))

    def slab_set_points(self, xyz_own, first_global, n_total, w=None):
        xyz_own = np.ascontiguousarray(xyz_own, np.float64)
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
        self.n = xyz_own.shape[0]
        self._check(self.lib.b200vor_slab_set_points(self.h, self.n, int(first_global), int(n_total),
                                                     _ptr(xyz_own), _ptr(w), 0))

    def slab_set_points_device(self, n_own, first_global, n_total, xyz_ptr, w_ptr=None):
        self.n = int(n_own)
        self._check(self.lib.b200vor_slab_set_points(self.h, self.n, int(first_global), int(n_total),
                                                     _ptr(int(xyz_ptr)), _ptr(int(w_ptr)) if w_ptr else None, 1))

    def slab_compute_cells(self, want_faces=False):
        p = CellParams(0, 0, int(want_faces), 0, 0)
        self._check(self.lib.b200vor_slab_compute_cells(self.h, ctypes.byref(p)))

    def slab_lloyd(self, iters, out=None, want_xyz=True):
        p = CellParams(0, 0, 0, 0, 0)
        if out is None and want_xyz:
            out = np.zeros((self.n, 3))
        self._check(self.lib.b200vor_slab_lloyd(self.h, int(iters), ctypes.byref(p), _ptr(out), 0))
        return out

    def slab_lloyd_device(self, iters, xyz_out_ptr=None):
        p = CellParams(0, 0, 0, 0, 0)
        self._check(self.lib.b200vor_slab_lloyd(self.h, int(iters), ctypes.byref(p),
                                                _ptr(int(xyz_out_ptr)) if xyz_out_ptr else None, 1))

    def slab_stats(self):
        s = SlabStats()
        self._check(self.lib.b200vor_slab_get_stats(self.h, ctypes.byref(s)))
        return {name: getattr(s, name) for name, _ in s._fields_}

    def exact_signs(self, planes):
        """planes: (n, 4, 4) float64 -- rows P0, P1, P2, P; returns the exact signs computed on the device"""
        planes = np.ascontiguousarray(planes, np.float64)
        out = np.zeros(planes.shape[0], np.int32)
        self._check(self.lib.b200vor_selftest_exact_sign(self.h, planes.shape[0], _ptr(planes), _ptr(out)))
        return out

    def microbench(self):
        """Measured FP64 / FP32 FMA throughput and the issue-slot peak of this GPU (dict)."""
        p = Peaks()
        self._check(self.lib.b200vor_microbench(self.h, ctypes.byref(p)))
        return {name: getattr(p, name) for name, _ in p._fields_ if name != "reserved"}

    def stats(self):
        s = Stats()
        self._check(self.lib.b200vor_get_stats(self.h, ctypes.byref(s)))
        return s.as_dict()
