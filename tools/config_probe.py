#!/usr/bin/env python
"""BASELINE configs 3 (Laguerre / OTM gradient) and 4 (clustered seeds, security-radius
retries) at size on one GPU: per-status cell counts, passes, conservation, device times.
usage: config_probe.py laguerre|clustered N [lloyd_iters] [a]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracles  # noqa: E402  (seed generators only)
from geogram_b200 import Context  # noqa: E402

NAMES = {1: "uncertified", 2: "uncertain", 4: "empty", 8: "boxface", 16: "overflow", 32: "duplicate",
         64: "degenerate"}


def status_hist(st):
    out = {"ok": int((st == 0).sum())}
    for bit, nm in NAMES.items():
        out[nm] = int(((st & bit) != 0).sum())
    return out


def main():
    kind, n = sys.argv[1], int(sys.argv[2])
    ctx = Context(0)
    res = {"config": kind, "n": n}
    if kind == "laguerre":
        iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
        a = float(sys.argv[4]) if len(sys.argv) > 4 else 0.25
        xyz = oracles.uniform_seeds(n, 12345)
        ctx.set_domain_periodic(1.0)
        ctx.set_points(xyz)
        xyz = ctx.lloyd(iters)                      # relax first
        h = n ** (-1.0 / 3.0)
        w = a * h * h * np.random.default_rng(5).random(n)
        t0 = time.perf_counter()
        ctx.set_points(xyz, w)
        ctx.compute_cells()
        res["wall_s_set_points_plus_cells"] = time.perf_counter() - t0
        m, mg, st = ctx.get_cells()
        s = ctx.stats()
        grad = m - 1.0 / n                          # OTM gradient g_i = vol_i - nu_i
        res.update(lloyd_iters=iters, a=a, status=status_hist(st), sum_m_minus_1=float(m.sum() - 1.0),
                   grad_sum=float(grad.sum()), grad_norm2=float(np.sqrt((grad * grad).sum())),
                   ms_grid=s["ms_grid"], ms_knn=s["ms_knn"], ms_clip=s["ms_clip"],
                   cells_per_s=n / (1e-3 * s["ms_total"]), passes=s["n_pass"][: s["passes"]],
                   k_pass=s["k_pass"][: s["passes"]])
    elif kind == "otm":
        # config 3, full solve: uniform target on relaxed seeds, random start weights
        iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
        eps = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-3
        xyz = oracles.uniform_seeds(n, 12345)
        ctx.set_domain_periodic(1.0)
        ctx.set_points(xyz)
        xyz = ctx.lloyd(iters)
        h = n ** (-1.0 / 3.0)
        w0 = 0.25 * h * h * np.random.default_rng(5).random(n)
        ctx.set_points(xyz)
        t0 = time.perf_counter()
        w, r = ctx.otm_newton(None, w0=w0, epsilon=eps)
        res["wall_s_otm_newton"] = time.perf_counter() - t0
        s = ctx.stats()
        res.update(r)
        res.update(ms_grid=s["ms_grid"], ms_clip=s["ms_clip"], ms_knn=s["ms_knn"])
    else:
        kmax = int(sys.argv[3]) if len(sys.argv) > 3 else 0
        mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
        xyz = oracles.clustered_seeds(n, 12345)
        t0 = time.perf_counter()
        ctx.set_points(xyz)
        ctx.compute_cells(kmax=kmax, mode=mode)
        res["wall_s_set_points_plus_cells"] = time.perf_counter() - t0
        m, mg, st = ctx.get_cells()
        s = ctx.stats()
        res.update(status=status_hist(st), sum_m_minus_1=float(m.sum() - 1.0),
                   ms_grid=s["ms_grid"], ms_knn=s["ms_knn"], ms_clip=s["ms_clip"],
                   cells_per_s=n / (1e-3 * s["ms_total"]), passes=s["n_pass"][: s["passes"]],
                   k_pass=s["k_pass"][: s["passes"]], grid=s["grid_dim"])
    print(json.dumps(res))


if __name__ == "__main__":
    main()
