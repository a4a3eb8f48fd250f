#!/usr/bin/env python
"""Multi-rank check of the slab + halo path (run under torchrun, one rank per GPU): every rank owns an
x-slab, computes its cells from own + ghost seeds (b200vor_slab_*), and compares them with the cells
of a single-context run over ALL seeds on its own GPU: face lists identical (global ids, periodic
instances), volumes to rounding, Lloyd iterates equal; periodic and box domains, with and without
weights.  Prints "slab check ok" on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geogram_b200 import Context  # noqa: E402
from geogram_b200.sharding import shard_range  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(os.environ.get("SLAB_N", "200000"))
    rng = np.random.default_rng(11)
    xyz = rng.random((n, 3))
    xyz = xyz[np.argsort(xyz[:, 0], kind="stable")]          # x-slabs
    h = n ** (-1.0 / 3.0)
    wts = 0.25 * h * h * rng.random(n)
    first, last = shard_range(n, rank, world)
    worst = 0.0
    for periodic in (True, False):
        for w in (None, wts):
            ref = Context(local)
            if periodic:
                ref.set_domain_periodic(1.0)
            ref.set_points(xyz, w)
            ref.compute_cells(mode=3, want_faces=True)
            m0, mg0, st0 = ref.get_cells()
            nf0, fn0, fa0 = ref.get_faces()
            ctx = Context(local)
            if periodic:
                ctx.set_domain_periodic(1.0)
            if world > 1:
                uid = [Context.comm_unique_id() if rank == 0 else None]   # one id per communicator
                dist.broadcast_object_list(uid, src=0)
                ctx.comm_init(uid[0], rank, world)
            ctx.slab_set_points(xyz[first:last], first, n, None if w is None else w[first:last])
            ctx.slab_compute_cells(want_faces=True)
            s = ctx.slab_stats()
            m, mg, st = ctx.get_cells()
            nf, fn, fa = ctx.get_faces()
            sl = slice(first, last)
            assert (st == st0[sl]).all() and ((st & ~np.uint8(4)) == 0).all(), (np.bincount(st), np.bincount(st0[sl]))
            assert (nf == nf0[sl]).all()
            live = np.arange(64)[None, :] < nf[:, None]
            assert (fn[live] == fn0[sl][live]).all(), "face ids differ"
            assert np.abs(fa[live] - fa0[sl][live]).max() < 1e-12
            err = np.abs(m - m0[sl]).max() / m0.max()
            assert err < 1e-12, err
            assert np.abs(mg - mg0[sl]).max() <= 1e-12 * m0.max()
            worst = max(worst, err)
            if world > 1:
                assert s["n_ghost_x"] > 0 and s["bytes_sent"] > 0, s
            if rank == 0:
                print("periodic", periodic, "weighted", w is not None, "rank0 stats", s, flush=True)
            # Lloyd (unweighted): same iterates as the single context
            if w is None:
                xa = ref.lloyd(2, mode=3)
                xb = ctx.slab_lloyd(2)
                d = np.abs(xa[sl] - xb).max()
                assert d < 1e-12, d
            ctx.close()
            ref.close()
    if world > 1:
        dist.barrier()
    if rank == 0:
        print("slab check ok: world %d, n %d, worst volume error %.2e" % (world, n, worst), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
