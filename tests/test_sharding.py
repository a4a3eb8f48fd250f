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
