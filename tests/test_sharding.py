"""world_size-2 (and 3, ragged) gloo test of the host-side sharding logic used by bench.py at
N > 1: ranges partition the seeds, the all-gather reassembles the updated rows identically on
every rank.  CPU only: the per-rank cell computation is replaced by a deterministic stand-in."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geogram_b200.sharding import allgather_rows, shard_range, shard_sizes  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    full = torch.rand((n_total, 3), dtype=torch.float64, generator=g)   # same seeds on every rank
    expect = full * 0.5 + 0.25                                          # the "Lloyd update" of all rows
    first, last = shard_range(n_total, rank, world)
    mine = full[first:last] * 0.5 + 0.25                                # this rank updates only its rows
    for _ in range(2):                                                  # two iterations, like bench.py
        allgather_rows(dist, full, mine)
        ok = bool((full == expect).all())
        mine = full[first:last].clone()
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    # device-timing pattern of bench.py: max over ranks
    ms = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((float(t.item()), float(ms.item())))
    dist.destroy_process_group()


def test_shard_ranges_partition():
    for n, w in [(10, 1), (10, 2), (10, 3), (12500000 * 8, 8), (7, 8)]:
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1


def test_slab_order_gives_spatial_slabs():
    """slab_order + shard_range: rank r owns the r-th x-quantile slab, equal counts"""
    import torch
    from geogram_b200.sharding import slab_order
    g = torch.Generator().manual_seed(5)
    full = torch.rand((1001, 3), dtype=torch.float64, generator=g)
    order = slab_order(full)
    assert sorted(order.tolist()) == list(range(1001))
    x = full[order][:, 0]
    prev_max = -1.0
    for r in range(4):
        a, b = shard_range(1001, r, 4)
        assert x[a:b].min().item() >= prev_max
        prev_max = x[a:b].max().item()


@pytest.mark.parametrize("world,n_total", [(2, 1000), (3, 1001)])
def test_allgather_rows_gloo(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok == 1.0
    assert ms == float(world)


class _StandInContext(object):
    """Host stand-in for geogram_b200.Context (same methods ShardedLloyd calls, raw addresses):
    the 'Lloyd update' of a row is x <- 0.5 x + 0.25 * mean(all seeds), so a rank that missed
    another rank's rows in the exchange gets a different answer."""

    def __init__(self):
        self.first, self.last = 0, None
        self.calls = []

    @staticmethod
    def _view(ptr, n):
        import ctypes
        import numpy as np
        return np.ctypeslib.as_array((ctypes.c_double * (3 * n)).from_address(ptr)).reshape(n, 3)

    def set_stream(self, s):
        raise AssertionError("no stream on the CPU path")

    def set_query_range(self, first, last):
        self.first, self.last = first, last

    def set_points_device(self, n, ptr):
        self.n = n
        self.xyz = self._view(ptr, n).copy()
        self.calls.append("set_points")

    def lloyd_device(self, iters, out_ptr=None, k0=0, mode=0):
        last = self.n if self.last is None else self.last
        for _ in range(iters):
            self.xyz[self.first:last] = 0.5 * self.xyz[self.first:last] + 0.25 * self.xyz.mean(axis=0)
        if out_ptr:
            self._view(out_ptr, self.n)[:] = self.xyz
        self.calls.append("lloyd")

    def stats(self):
        return {"ms_grid": 1.0, "ms_knn": 0.0, "ms_clip": 2.0}


def _lloyd_worker(rank, world, port, n_total, q):
    from geogram_b200.sharding import ShardedLloyd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(11)
    seeds = torch.rand((n_total, 3), dtype=torch.float64, generator=g)
    # single-process answer
    ref = seeds.clone()
    for _ in range(3):
        ref = 0.5 * ref + 0.25 * ref.mean(dim=0)
    job = ShardedLloyd(_StandInContext(), n_total, dist, torch.device("cpu"))
    first, last = job.first, job.last
    job.set_shard_host(seeds[first:last].clone())      # own rows in, the rest over the wire
    for _ in range(3):
        job.step()
    mine = torch.empty((last - first, 3), dtype=torch.float64)
    job.get_shard_host(mine)
    err = float((mine - ref[first:last]).abs().max())
    full_err = float((job.full - ref).abs().max())
    t = torch.tensor([max(err, full_err)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((float(t.item()), job.phase, job.ctx.calls))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 600), (3, 1001)])
def test_sharded_lloyd_gloo(world, n_total):
    """ShardedLloyd (the multi-GPU Lloyd driver of bench.py) on world_size 2 and 3 with gloo:
    shard upload, exchange, per-rank update of the own rows, re-gather, read-back."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lloyd_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, phase, calls = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-14
    # the grid of the re-gathered seeds is built lazily, by the step that needs it: no rebuild
    # after the last step (a caller that uploads fresh seeds every step pays for one build, not two)
    assert calls == ["set_points", "lloyd"] + ["set_points", "lloyd"] * 2
    assert phase["ms_clip"] == 6.0 and phase["ms_grid"] == 3.0


# ---- slab + halo (csrc/slab.cu): the host statement of the halo selection, exchanged over gloo, must
#      give every rank exactly the seeds its own cells need: the cells the CPU oracle computes from
#      the rank's LOCAL set (a plain box problem) are the cells of the periodic run on all seeds ----
def _slab_worker(rank, world, port, n_total, periodic, q):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracles
    from geogram_b200.sharding import halo_local_set, halo_send_list, slab_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyz = np.random.default_rng(3).random((n_total, 3))
    xyz = xyz[np.argsort(xyz[:, 0], kind="stable")]
    # a few seeds wrapped around the period by a Lloyd step (rank 0 owns seeds next to x = 1)
    if periodic:
        xyz[:3, 0] = 1.0 - 1e-3 * (1 + np.arange(3))
    first, last = shard_range(n_total, rank, world)
    own = xyz[first:last]
    period = 1.0 if periodic else None
    H = 3.0 * (8.0 / n_total) ** (1.0 / 3.0)
    bounds = [None] * world
    dist.all_gather_object(bounds, slab_bounds(own[:, 0], period))
    send = halo_send_list(own, first, rank, bounds, H, period)
    inbox = [None] * world
    for p in range(world):                 # peer p collects what every rank has for it
        got = [None] * world
        dist.all_gather_object(got, send[p])
        if p == rank:
            inbox = got
    lx, lgid, lt = halo_local_set(own, first, rank, bounds, inbox, H, period)
    lo, hi, _ = bounds[rank]
    port_ = oracles.Port()
    if periodic:
        box = [lo - H, -H, -H, hi + H, 1.0 + H, 1.0 + H]
    else:
        box = [max(lo - H, 0.0), 0.0, 0.0, min(hi + H, 1.0), 1.0, 1.0]
    L = port_.cells(lx, box=box)
    G = port_.cells(xyz, periodic=periodic)
    n_own = last - first
    err = float(np.abs(L.m[:n_own] - G.m[first:last]).max() / G.m.max())
    # same neighbours: local face ids -> global id + instance
    bad = 0
    for i in range(0, n_own, 7):
        ids = L.face_nbr[i, : L.nfaces[i]]
        loc = ids[ids < oracles.BOX_ID]
        if periodic and len(loc) != len(ids):
            bad += 1                      # a periodic cell never touches the local box
        rel = lt[loc] - lt[i]
        inst = np.array([oracles.translation_to_instance(*r) for r in rel], np.int64)
        mine = np.sort(lgid[loc] + inst * n_total)
        ref = G.face_nbr[first + i, : G.nfaces[first + i]]
        ref = np.sort(ref[ref < oracles.BOX_ID].astype(np.int64))
        bad += 0 if (len(mine) == len(ref) and (mine == ref).all()) else 1
    t = torch.tensor([err, float(bad), float(len(lx))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((float(t[0]), float(t[1]), float(t[2])))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic", [(2, True), (3, True), (2, False)])
def test_slab_halo_plan_gloo(world, periodic):
    n_total = 6000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, n_total, periodic, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, bad, nlocal = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12 and bad == 0
