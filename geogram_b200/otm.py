"""Multilevel driver of the semi-discrete OT solve above the C ABI (b200vor_otm_newton,
b200vor_knn_points): the Python statement of B200GEO::OptimalTransportMap3d::optimize_levels
(geogram_b200/host/b200geo.cpp).  Reference call site: OptimalTransportMap3d::optimize_levels,
src/examples/exploragram/compute_OTM/main.cpp:292-297 (exploragram itself is not in the reference
tree: the published scheme -- solve the prefixes of a BRIO-ordered point set, extrapolate the
weights to the points that enter -- is restated, parity unpinned)."""
import numpy as np

from .capi import B200VorError


def extrapolate_weights(pts, w, q, rows):
    """Start weights of the new points q (m, 3): value at q of the least-squares plane through
    (pts[rows[i]], w[rows[i]]), clamped to the range of the fitted weights (no overshoot outside the
    hull of the neighbours); a rank-deficient neighbourhood takes the nearest weight."""
    m, k = rows.shape
    if k < 4:
        return w[rows[:, 0]].copy()
    d = pts[rows] - q[:, None, :]                      # (m, k, 3)
    f = np.concatenate([np.ones((m, k, 1)), d], axis=2)  # (m, k, 4)
    wk = w[rows]                                       # (m, k)
    A = np.einsum("mki,mkj->mij", f, f)
    b = np.einsum("mki,mk->mi", f, wk)
    out = wk[:, 0].copy()
    scale = A[:, 1, 1] + A[:, 2, 2] + A[:, 3, 3]
    ok = np.linalg.cond(A) < 1e10
    ok &= scale > 0
    if ok.any():
        x = np.linalg.solve(A[ok], b[ok][:, :, None])[:, 0, 0]
        out[ok] = np.minimum(np.maximum(x, wk[ok].min(axis=1)), wk[ok].max(axis=1))
    return out


def optimize_levels(ctx, xyz, levels, nu=None, epsilon=0.01, max_newton=0, period=0.0, k_fit=24, log=None):
    """ctx: a Context whose domain is set; xyz (n, 3) in BRIO order; levels = [0, e_1, ..., n].
    Returns (w, list of per-level result dicts)."""
    xyz = np.ascontiguousarray(xyz, np.float64)
    n = xyz.shape[0]
    if len(levels) < 2 or levels[0] != 0 or levels[-1] != n or any(b >= e for b, e in zip(levels, levels[1:])):
        raise ValueError("levels must increase from 0 to the number of points")
    w = np.zeros(n)
    results = []
    for b, e in zip(levels, levels[1:]):
        if b:
            # the context still holds the points [0, b): one batched by-point kNN query (a box query)
            if period > 0.0:
                ctx.set_domain_box([0, 0, 0, period, period, period])
            rows, _ = ctx.knn_points(xyz[:e], min(b, k_fit))
            if period > 0.0:
                ctx.set_domain_periodic(period)
            # ALL points of the new level, the solved ones included, start from the fitted (low-pass)
            # field: what carries over between levels is the smooth part of the potential; the per-point
            # part compensates the cell irregularities of ONE level's point set (order h_coarse^2, as large
            # as the squared spacing of the next level) and empties cells when it is carried along
            w[:e] = extrapolate_weights(xyz[:b], w[:b].copy(), xyz[:e], rows.astype(np.int64))
        nu_l = None
        if nu is not None:
            nu_l = np.asarray(nu[:e], np.float64) * (np.sum(nu) / np.sum(nu[:e]))
        ctx.set_points(xyz[:e])
        for attempt in range(9):
            try:
                we, r = ctx.otm_newton(nu=nu_l, w0=w[:e], epsilon=epsilon, max_newton=max_newton)
                break
            except B200VorError as err:
                msg = str(err)
                if attempt == 8 or not ("an initial cell is empty" in msg or "of the start weights are flagged" in msg):
                    raise
                if log:
                    log("level %d..%d: %s -> weights halved" % (b, e, msg))
                w[:e] *= 0.5      # towards the Voronoi diagram, whose cells are all non-empty
        w[:e] = we
        r["points"] = e
        r["halvings"] = attempt
        results.append(r)
    return w, results
