"""GPU parity of the thread-per-cell kernel (tpc.cu): the default first pass of
b200vor_compute_cells / b200vor_lloyd when no face lists are requested.  Checked against the
CPU oracle (volumes, centroids), against the team kernel on the same input (mode 3), and for
the size-independent properties at a BASELINE-scale input."""
import numpy as np
import pytest

import oracles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Context():
    from geogram_b200 import Context as C
    return C


@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("n", [20000, 100000])
def test_tpc_vs_oracle_and_team_kernel(Context, port, n, periodic):
    xyz = oracles.uniform_seeds(n, 77 + n)
    ctx = Context(0)
    if periodic:
        ctx.set_domain_periodic(1.0)
    ctx.set_points(xyz)
    ctx.compute_cells(mode=4)                # thread-per-cell kernel first
    s = ctx.stats()
    assert s["tpc_cells"] == n               # the new kernel really ran on every cell
    m, mg, st = ctx.get_cells()
    assert (st == 0).all(), np.bincount(st)
    O = port.cells(xyz, periodic=periodic)
    assert np.all(np.abs(m - O.m) <= 1e-9 * O.m)
    assert np.all(np.abs(mg - O.mg) <= 1e-9 * O.m[:, None])
    assert abs(m.sum() - 1.0) < 1e-9
    # event counters of the SURVEY 8(d) flop model are those of the reference algorithm
    assert 0.8 * n * 14 < s["sum_planes"] < n * 60
    # same cells as the team kernel
    ctx.compute_cells(mode=3)
    assert ctx.stats()["tpc_cells"] == 0
    m3, mg3, st3 = ctx.get_cells()
    assert (st3 == 0).all()
    assert np.all(np.abs(m - m3) <= 1e-11 * m3)
    assert np.all(np.abs(mg - mg3) <= 1e-11 * m3[:, None])


def test_tpc_flagged_cells_are_redone(Context, port):
    """Cells the thread-per-cell kernel cannot finish inside its capacities are handed to the
    team kernel: every cell ends OK and the pass list shows the hand-over."""
    n = 200000
    xyz = oracles.uniform_seeds(n, 4242)
    ctx = Context(0)
    ctx.set_domain_periodic(1.0)
    ctx.set_points(xyz)
    ctx.compute_cells(mode=4)
    s = ctx.stats()
    m, mg, st = ctx.get_cells()
    assert (st == 0).all(), np.bincount(st)
    assert s["passes"] >= 2 and s["n_pass"][0] == n and 0 < s["n_pass"][1] < 0.1 * n
    assert abs(m.sum() - 1.0) < 1e-9
    O = port.cells(xyz, periodic=True)
    assert np.all(np.abs(m - O.m) <= 1e-9 * O.m)


def test_tpc_lloyd_matches_team_kernel(Context):
    n = 50000
    xyz = oracles.uniform_seeds(n, 9)
    a = Context(0)
    a.set_domain_periodic(1.0)
    a.set_points(xyz)
    xa = a.lloyd(3, mode=4)
    b = Context(0)
    b.set_domain_periodic(1.0)
    b.set_points(xyz)
    xb = b.lloyd(3, mode=3)
    assert np.abs(xa - xb).max() < 1e-12


def test_tpc_duplicates_and_small_sets(Context, port):
    xyz = oracles.uniform_seeds(3000, 3)
    xyz[10] = xyz[500]                       # exact duplicate: the higher index gets no cell
    ctx = Context(0)
    ctx.set_points(xyz)
    ctx.compute_cells(mode=4)
    m, mg, st = ctx.get_cells()
    assert st[500] & 32 and m[500] == 0.0 and st[10] == 0
    assert abs(m.sum() - 1.0) < 1e-9
    for n in (1, 2, 5, 40):
        p = oracles.uniform_seeds(n, n)
        ctx.set_points(p)
        ctx.compute_cells(mode=4)
        m, mg, st = ctx.get_cells()
        assert (st == 0).all() and abs(m.sum() - 1.0) < 1e-9


def test_tpc_conservation_10m(Context):
    """BASELINE configs[1] size: all cells certified, total volume conserved to 1e-9."""
    import torch
    n = 10000000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(12345)
    x = torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=gen)
    ctx = Context(0)
    ctx.set_domain_periodic(1.0)
    ctx.set_points_device(n, x.data_ptr())
    ctx.compute_cells(mode=4)
    m = torch.empty((n,), dtype=torch.float64, device="cuda")
    st = torch.empty((n,), dtype=torch.uint8, device="cuda")
    ctx.get_cells_device(m.data_ptr(), None, st.data_ptr())
    assert int((st != 0).sum().item()) == 0
    assert abs(m.sum().item() - 1.0) < 1e-9
    assert ctx.stats()["tpc_cells"] == n


@pytest.mark.parametrize("periodic", [True, False])
def test_tpc_face_lists_vs_oracle(Context, port, periodic):
    n = 100000
    xyz = oracles.uniform_seeds(n, 31)
    ctx = Context(0)
    if periodic:
        ctx.set_domain_periodic(1.0)
    ctx.set_points(xyz)
    ctx.compute_cells(mode=4, want_faces=True)
    assert ctx.stats()["tpc_cells"] == n
    m, mg, st = ctx.get_cells()
    nf, fn, fa = ctx.get_faces()
    assert (st == 0).all()
    O = port.cells(xyz, periodic=periodic, areas=True)
    col = np.arange(64)[None, :]
    assert (nf == O.nfaces).all()
    assert not ((fn != O.face_nbr) & (col < O.nfaces[:, None])).any()      # connectivity bit-exact
    mask = col < O.nfaces[:, None]
    assert np.abs(fa[mask] - O.face_area[mask]).max() < 1e-12
    assert np.all(np.abs(m - O.m) <= 1e-9 * O.m)


@pytest.mark.parametrize("mode", [0, 3, 4])
def test_degenerate_lattice_resolved_exactly(Context, mode):
    """A cubic lattice (every sign test of the far planes is an exact tie): whatever kernel runs
    first, the flagged cells are re-done with exact sign decisions and the reference's default tie
    rule (convex_cell.cpp:905-911) -- status 0 everywhere, face lists and volumes equal to the
    fixture the unmodified reference produced in exact mode."""
    from conftest import Golden
    G = Golden("cells_box_lattice_exact_1728.npz")
    ctx = Context(0)
    ctx.set_domain_box([0, 0, 0, 0.75, 0.75, 0.75])
    ctx.set_points(G.xyz)
    ctx.compute_cells(mode=mode, want_faces=True)
    s = ctx.stats()
    m, mg, st = ctx.get_cells()
    nf, fn, fa = ctx.get_faces()
    assert (st == 0).all(), np.bincount(st)
    assert s["exact_cells"] > 0 or mode == 4
    assert (nf == G.nfaces).all() and (fn == G.face_nbr).all()
    assert np.abs(m - G.m).max() < 1e-15
    assert abs(m.sum() - 0.75 ** 3) < 1e-14
    assert np.abs(fa - G.face_area).max() < 1e-14


def test_exact_predicate_on_device_vs_rational_arithmetic(Context):
    """The device's exact sign predicate (exact_pred.cuh through b200vor_selftest_exact_sign) against
    python Fractions: random, nearly degenerate and exactly degenerate (dyadic) plane quadruples."""
    from fractions import Fraction as Fr

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

    rng = np.random.default_rng(3)
    cases = []
    for it in range(600):
        Pl = rng.standard_normal((4, 4))
        if it % 3 == 0:
            v = np.linalg.solve(Pl[:3, :3], -Pl[:3, 3])
            Pl[3, 3] = -(Pl[3, :3] @ v)
        if it % 3 == 1:
            Pl = np.round(Pl * 8) / 8
            if abs(np.linalg.det(Pl[:3, :3])) < 1e-9:
                continue
            v = np.round(np.linalg.solve(Pl[:3, :3], -Pl[:3, 3]) * 64) / 64
            for k in range(4):
                Pl[k, 3] = -(Pl[k, :3] @ v)
        cases.append(Pl)
    cases = np.array(cases)
    got = Context(0).exact_signs(cases)
    want = np.array([exact(c) for c in cases])
    assert (got == want).all()
    assert (want == 0).sum() > 150
