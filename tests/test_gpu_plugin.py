"""The geogram plugin (geogram_b200/plugin): real GEO::NearestNeighborSearch / GEO::Delaunay
subclasses registered with the reference's own factories, compiled against the reference headers
and linked with the UNMODIFIED reference (oracle/_ref/libgeoref.so) -- a test target, prebuilt in
the build container by `make -C oracle plugin` and shipped in oracle/_ref/ (the GPU box has no
/root/reference).  The acceptance program restates the reference's own test_nn_search criterion
(src/tests/test_nn_search/main.cpp:136-181) and drives Delaunay::create(3, "B200")."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracles

pytestmark = pytest.mark.gpu

EXE = os.path.join(oracles.ROOT, "oracle", "_ref", "test_plugin")


@pytest.fixture(scope="module")
def run(tmp_path_factory):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/test_plugin not built (needs /root/reference: make -C oracle plugin)")
    d = tmp_path_factory.mktemp("plugin")
    obj, pts = str(d / "cell.obj"), str(d / "pts.bin")
    r = subprocess.run([EXE, "20000", "21", obj, pts], capture_output=True, text=True, timeout=600)
    return r, obj, pts


def test_plugin_acceptance_program(run):
    r, _, _ = run
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "ALL CHECKS PASSED" in r.stdout
    checks = re.findall(r"check (\S+) ok", r.stdout)
    for name in ("nn_by_index_sq_dist_bit_equal", "nn_by_pointer_sq_dist_bit_equal",
                 "nn_arbitrary_points_parallel_for_bit_equal", "nn_point_queries_are_batched",
                 "delaunay_neighbors_identical_to_NN", "delaunay_enlarge_neighborhood",
                 "convex_cell_volume_matches_engine", "convex_cell_faces_match_reference_clipping",
                 "convex_cell_append_to_mesh", "convex_cell_further_clip"):
        assert name in checks
    m = re.search(r"info by-point queries (\d+) launches (\d+) threads (\d+)", r.stdout)
    nq, launches, threads = (int(x) for x in m.groups())
    if threads > 1:
        assert launches < nq          # concurrent callers were combined into shared launches


def test_save_byte_equal_to_ConvexCell_save(run):
    """The OBJ text of geogram_b200.export.save_cell_obj == what the reference's ConvexCell::save
    (voronoi/convex_cell.cpp:282-352) wrote for the same cell, byte for byte."""
    from geogram_b200 import Context, export
    r, obj, pts = run
    pick = int(re.search(r"info saved cell (\d+)", r.stdout).group(1))
    xyz = np.fromfile(pts, np.float64).reshape(-1, 3)
    ctx = Context(0)
    ctx.set_points(xyz)
    ctx.compute_cells(want_faces=2)
    nf, fn, fa = ctx.get_faces()
    nv, vx, vf = ctx.get_vertices()
    mine = export.save_cell_obj(int(nv[pick]), vx[pick], vf[pick], int(nf[pick]))
    ref = open(obj).read()
    assert mine == ref
