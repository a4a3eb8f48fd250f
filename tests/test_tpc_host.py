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

    def run(xyz, periodic, box=oracles.UNIT_BOX, ppc=8.0, target=20.0):
        n = xyz.shape[0]
        m, mg = np.zeros(n), np.zeros((n, 3))
        st, ev = np.zeros(n, np.uint8), np.zeros((n, 2), np.uint32)
        box = np.ascontiguousarray(box, np.float64)
        vp = ctypes.c_void_p
        lib.tpc_host_cells(ctypes.c_uint32(n), xyz.ctypes.data_as(vp), int(periodic), ctypes.c_double(1.0),
                           box.ctypes.data_as(vp), ctypes.c_double(ppc), ctypes.c_float(target),
                           m.ctypes.data_as(vp), mg.ctypes.data_as(vp), st.ctypes.data_as(vp), ev.ctypes.data_as(vp))
        return m, mg, st, ev
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
