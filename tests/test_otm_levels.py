"""Multilevel OT driver (geogram_b200/otm.py; reference call site
src/examples/exploragram/compute_OTM/main.cpp:292-297, OptimalTransportMap3d::optimize_levels)."""
import numpy as np
import pytest

from geogram_b200.otm import extrapolate_weights, optimize_levels


def _rows(pts, q, k):
    d = ((pts[None] - q[:, None]) ** 2).sum(2)
    return np.argsort(d, axis=1, kind="stable")[:, :k]


def test_extrapolated_weights_reproduce_a_linear_field_and_never_overshoot():
    rng = np.random.default_rng(3)
    pts = rng.random((200, 3))
    w = 0.3 - 1.5 * pts[:, 0] + 0.25 * pts[:, 1] + 2.0 * pts[:, 2]
    q = rng.random((64, 3)) * 1.4 - 0.2          # some queries outside the hull of the points
    rows = _rows(pts, q, 12)
    v = extrapolate_weights(pts, w, q, rows)
    exact = 0.3 - 1.5 * q[:, 0] + 0.25 * q[:, 1] + 2.0 * q[:, 2]
    lo, hi = w[rows].min(axis=1), w[rows].max(axis=1)
    assert np.abs(v - np.clip(exact, lo, hi)).max() < 1e-12
    inside = (exact >= lo) & (exact <= hi)
    assert inside.sum() > 20 and np.abs(v[inside] - exact[inside]).max() < 1e-12
    # rank-deficient neighbourhoods: nearest weight
    flat = pts.copy()
    flat[:, 2] = 0.5
    assert np.array_equal(extrapolate_weights(flat, w, q, rows), w[rows[:, 0]])
    assert np.array_equal(extrapolate_weights(pts, w, q, rows[:, :3]), w[rows[:, 0]])


def test_levels_are_validated():
    with pytest.raises(ValueError):
        optimize_levels(None, np.zeros((10, 3)), [0, 5, 5, 10])
    with pytest.raises(ValueError):
        optimize_levels(None, np.zeros((10, 3)), [1, 10])


@pytest.mark.gpu
@pytest.mark.parametrize("periodic", [False, True])
def test_multilevel_solve_equals_single_level(periodic):
    from geogram_b200 import Context
    n = 4000
    rng = np.random.default_rng(21)
    ctx = Context(0)
    if periodic:
        ctx.set_domain_periodic(1.0)
    else:
        ctx.set_domain_box([0, 0, 0, 1, 1, 1])
    ctx.set_points(rng.random((n, 3)))
    xyz = ctx.lloyd(5)
    nu = 1.0 + 0.1 * np.sin(2 * np.pi * xyz[:, 0])     # smooth target, weight range ~ 8 h^2 (certified regime)
    nu /= nu.sum()
    ctx.set_points(xyz)
    w1, r1 = ctx.otm_newton(nu=nu, epsilon=1e-4)
    w2, rs = optimize_levels(ctx, xyz, [0, n // 16, n // 4, n], nu=nu, epsilon=1e-4,
                             period=1.0 if periodic else 0.0)
    assert r1["converged"] and all(r["converged"] for r in rs) and len(rs) == 3
    # the same transport map: weights equal up to the free constant, volumes on target
    d = (w2 - w2.mean()) - (w1 - w1.mean())
    assert np.abs(d).max() <= 1e-2 * np.ptp(w1)
    ctx.set_points(xyz, w2)
    ctx.compute_cells()
    m, _, st = ctx.get_cells()
    assert (st == 0).all() and np.abs(m / nu - 1.0).max() <= 1e-4 and abs(m.sum() - 1.0) < 1e-9
    # every level is a short Newton run (the halving safeguard may act on a coarse level: a start, not a result)
    print("levels:", [(r["points"], r["newton_iters"], r["halvings"]) for r in rs], "single level:", r1["newton_iters"])
    assert all(r["newton_iters"] <= 25 for r in rs)
