"""The thread-per-cell kernel body (geogram_b200/csrc/tpc_cell.cuh, __host__ __device__) run on the
CPU by tests/host/tpc_host.cpp and checked against the oracle: the logic of the GPU kernel
(fp32-filtered clipping, shell walk, orthoscheme integration, capacity flags) without a GPU.
TEST INFRASTRUCTURE: the product never runs this code on the host."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracles

ROOT = oracles.ROOT
SO = os.path.join(ROOT, "tests", "_build", "libtpc_host.so")


@pytest.fixture(scope="module")
def tpc():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "host", "tpc_host.cpp"), "-o", SO])
    lib = ctypes.CDLL(SO)

    def run(xyz, periodic, box=oracles.UNIT_BOX, ppc=8.0, target=20.0, big=0, faces=False):
        n = xyz.shape[0]
        m, mg = np.zeros(n), np.zeros((n, 3))
        st, ev = np.zeros(n, np.uint8), np.zeros((n, 2), np.uint32)
        nf, fn, fa = np.zeros(n, np.uint32), np.zeros((n, 64), np.uint32), np.zeros((n, 64))
        box = np.ascontiguousarray(box, np.float64)
        vp = ctypes.c_void_p
        lib.tpc_host_cells_ex(ctypes.c_uint32(n), xyz.ctypes.data_as(vp), int(periodic), ctypes.c_double(1.0),
                              box.ctypes.data_as(vp), ctypes.c_double(ppc), ctypes.c_float(target),
                              m.ctypes.data_as(vp), mg.ctypes.data_as(vp), st.ctypes.data_as(vp), ev.ctypes.data_as(vp),
                              int(big), nf.ctypes.data_as(vp) if faces else None,
                              fn.ctypes.data_as(vp) if faces else None, fa.ctypes.data_as(vp) if faces else None)
        if faces:
            return m, mg, st, ev, nf, fn, fa
        return m, mg, st, ev
    run.lib = lib
    return run


@pytest.mark.parametrize("periodic", [True, False])
def test_tpc_body_vs_oracle(tpc, port, periodic):
    n = 20000
    xyz = oracles.uniform_seeds(n, 4242)
    m, mg, st, ev = tpc(xyz, periodic)
    O = port.cells(xyz, periodic=periodic)
    ok = st == 0
    # the only flags are the capacity ones (handed to the team kernel on the GPU), on few cells
    assert set(np.unique(st)) <= {0, 16, 17, 1}
    assert ok.mean() > 0.95
    assert (m[~ok] == 0).all()
    assert np.all(np.abs(m[ok] - O.m[ok]) <= 1e-9 * O.m[ok])
    assert np.all(np.abs(mg[ok] - O.mg[ok]) <= 1e-9 * O.m[ok][:, None])
    # event counts = those of the reference algorithm on the same input (P 43.5, C 906, V 68.2)
    nP, nV, nC = ev[ok, 0] & 0xffff, (ev[ok, 0] >> 16) & 0xfff, ev[ok, 1] & 0xfffff
    assert 38 < nP.mean() < 48 and 60 < nV.mean() < 76 and 800 < nC.mean() < 1000


def test_tpc_body_relaxed_and_duplicates(tpc, port):
    n = 8000
    xyz = oracles.uniform_seeds(n, 5)
    for _ in range(3):
        A = port.cells(xyz, periodic=True)
        xyz = port.lloyd_update(xyz, A.m, A.mg, True, 1.0)
    O = port.cells(xyz, periodic=True)
    m, mg, st, ev = tpc(xyz, True)
    ok = st == 0
    assert ok.mean() > 0.99
    assert np.all(np.abs(m[ok] - O.m[ok]) <= 1e-9 * O.m[ok])
    # relaxed cells are cheaper: fewer planes inside the security radius
    assert (ev[ok, 0] & 0xffff).mean() < 30
    # exact duplicate (delaunay_nn.cpp:122-135): the higher index gets no cell; its neighbours see
    # the same plane twice (exact ties) and are flagged UNCERTAIN, never silently wrong
    xyz = np.ascontiguousarray(xyz)
    xyz[4000] = xyz[17]
    m, mg, st, ev = tpc(xyz, True)
    assert st[4000] & 32 and m[4000] == 0.0 and not (st[17] & 32)
    assert set(np.unique(st)) <= {0, 1, 2, 16, 17, 18, 32, 34}


@pytest.mark.parametrize("periodic", [True, False])
def test_tpc_body_face_lists_and_areas(tpc, port, periodic):
    n = 12000
    xyz = oracles.uniform_seeds(n, 99)
    m, mg, st, ev, nf, fn, fa = tpc(xyz, periodic, faces=True)
    O = port.cells(xyz, periodic=periodic, areas=True)
    ok = st == 0
    col = np.arange(64)[None, :]
    bad = (O.nfaces != nf) | ((O.face_nbr != fn) & (col < O.nfaces[:, None])).any(axis=1)
    assert int((bad & ok).sum()) == 0          # connectivity bit-exact
    mask = (col < O.nfaces[:, None]) & ok[:, None]
    assert np.abs(fa[mask] - O.face_area[mask]).max() < 1e-12


def test_exact_predicate_vs_rational_arithmetic(tpc):
    """exact_pred.cuh: sign of a plane at the intersection of three planes, by expansions, against
    python Fractions -- random, nearly degenerate and exactly degenerate (dyadic) inputs."""
    from fractions import Fraction as Fr
    lib = tpc.lib

    def det3(a, b, c):
        return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0])

    def exact(Pl):
        F = [[Fr(float(x)) for x in p] for p in Pl]
        W = det3(F[0][:3], F[1][:3], F[2][:3])
        if W == 0:
            return 2
        N = (F[3][3] * W - F[0][3] * det3(F[3][:3], F[1][:3], F[2][:3]) - F[1][3] * det3(F[3][:3], F[2][:3], F[0][:3])
             - F[2][3] * det3(F[3][:3], F[0][:3], F[1][:3]))
        return ((N > 0) - (N < 0)) * ((W > 0) - (W < 0))

    rng = np.random.default_rng(1)
    zeros = 0
    for it in range(900):
        Pl = rng.standard_normal((4, 4))
        if it % 3 == 0:      # fourth plane through the vertex, up to rounding
            v = np.linalg.solve(Pl[:3, :3], -Pl[:3, 3])
            Pl[3, 3] = -(Pl[3, :3] @ v)
        if it % 3 == 1:      # dyadic: all four planes EXACTLY through one point
            Pl = np.round(Pl * 8) / 8
            if abs(np.linalg.det(Pl[:3, :3])) < 1e-9:
                continue
            v = np.round(np.linalg.solve(Pl[:3, :3], -Pl[:3, 3]) * 64) / 64
            for k in range(4):
                Pl[k, 3] = -(Pl[k, :3] @ v)
        a = [np.ascontiguousarray(x, np.float64) for x in Pl]
        got = lib.tpc_exact_sign(*[x.ctypes.data_as(ctypes.c_void_p) for x in a])
        want = exact(Pl)
        zeros += want == 0
        assert got == want
    assert zeros > 200


def test_tpc_body_lattice_exact_mode_vs_reference_golden(tpc):
    """Degenerate input (cubic lattice, dyadic coordinates): the large instance decides the ties
    exactly with the reference's default rule; face lists and volumes equal the golden fixture made
    by the unmodified reference in exact mode (tests/golden/make_golden.py lattice)."""
    from conftest import Golden
    G = Golden("cells_box_lattice_exact_1728.npz")
    box = np.array([0, 0, 0, 0.75, 0.75, 0.75])
    m, mg, st, ev, nf, fn, fa = tpc(G.xyz, False, box=box, big=1, faces=True)
    assert (st == 0).all()
    assert (nf == G.nfaces).all() and (fn == G.face_nbr).all()
    assert np.abs(m - G.m).max() < 1e-15 and abs(m.sum() - 0.75 ** 3) < 1e-14
    assert np.abs(fa - G.face_area).max() < 1e-14
    assert (nf == 26).sum() == 1000            # interior cells: 6 + 12 + 8 faces, 20 of zero area
