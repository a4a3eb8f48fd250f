"""GPU tests of the slab + halo path (slab.cu: b200vor_slab_*): the cells a rank computes from its own
seeds plus the ghosts of a halo must be the cells of the single-context run on all seeds -- same face
lists (global ids, periodic instances), volumes to rounding.  world = 1 exercises the whole ghost
machinery on one GPU (the ghosts are then the periodic images only); the 2-rank case needs two GPUs
and NCCL (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracles

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def Context():
    from geogram_b200 import Context as C
    return C


def _single(Context, xyz, periodic, w=None, faces=True):
    ctx = Context(0)
    if periodic:
        ctx.set_domain_periodic(1.0)
    ctx.set_points(xyz, w)
    ctx.compute_cells(mode=3, want_faces=faces)
    m, mg, st = ctx.get_cells()
    nf, fn, fa = ctx.get_faces() if faces else (None, None, None)
    return m, mg, st, nf, fn, fa


@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("weighted", [False, True])
def test_slab_world1_equals_single_context(Context, periodic, weighted):
    n = 60000
    xyz = oracles.uniform_seeds(n, 5)
    w = None
    if weighted:
        h = n ** (-1.0 / 3.0)
        w = 0.25 * h * h * np.random.default_rng(1).random(n)
    m0, mg0, st0, nf0, fn0, fa0 = _single(Context, xyz, periodic, w)
    ctx = Context(0)
    if periodic:
        ctx.set_domain_periodic(1.0)
    ctx.slab_set_points(xyz, 0, n, w)
    ctx.slab_compute_cells(want_faces=True)
    s = ctx.slab_stats()
    assert s["n_own"] == n and s["n_flagged"] == 0
    if periodic:
        assert s["n_ghost_x"] > 0 and s["n_ghost_yz"] > 0 and s["n_local"] == n + s["n_ghost_x"] + s["n_ghost_yz"]
    else:
        assert s["n_local"] == n
    m, mg, st = ctx.get_cells()
    nf, fn, fa = ctx.get_faces()
    assert m.shape[0] == n
    # (random weights empty a few Laguerre cells: status EMPTY, the same ones in both runs)
    assert (st == st0).all() and ((st & ~np.uint8(4)) == 0).all(), (np.bincount(st), np.bincount(st0))
    assert (nf == nf0).all()
    col = np.arange(64)[None, :]
    live = col < nf[:, None]
    assert (fn[live] == fn0[live]).all()            # global ids + n * instance, same order
    assert np.abs(fa[live] - fa0[live]).max() < 1e-12
    assert np.abs(m - m0).max() <= 1e-12 * m0.max()
    assert np.abs(mg - mg0).max() <= 1e-12 * m0.max()
    assert abs(m.sum() - 1.0) < 1e-11


def test_slab_world1_lloyd_equals_single_context(Context):
    n = 80000
    xyz = oracles.uniform_seeds(n, 6)
    a = Context(0)
    a.set_domain_periodic(1.0)
    a.set_points(xyz)
    xa = a.lloyd(3, mode=3)
    b = Context(0)
    b.set_domain_periodic(1.0)
    b.slab_set_points(xyz, 0, n)
    xb = b.slab_lloyd(3)
    assert np.abs(xa - xb).max() < 1e-12
    s = b.slab_stats()
    assert s["retries"] == 0 and s["n_flagged"] == 0
    # a slab context refuses the single-context calls (its point array holds ghosts) until set_points
    from geogram_b200 import B200VorError
    for call in (lambda: b.compute_cells(), lambda: b.lloyd(1), lambda: b.knn(4)):
        with pytest.raises(B200VorError):
            call()
    b.set_points(xyz)
    b.compute_cells()
    assert (b.get_cells()[2] == 0).all()


def test_slab_requeues_with_a_wider_halo(Context):
    """A halo of a fraction of a bin cannot hold the security radius: the pass must double it until
    no cell leaves it, and still return the right cells."""
    n = 30000
    xyz = oracles.uniform_seeds(n, 8)
    m0, mg0, st0, _, _, _ = _single(Context, xyz, True, faces=False)
    ctx = Context(0)
    ctx.set_domain_periodic(1.0)
    ctx.slab_set_halo_bins(0.4)
    ctx.slab_set_points(xyz, 0, n)
    ctx.slab_compute_cells()
    s = ctx.slab_stats()
    assert s["retries"] >= 1 and s["n_flagged"] == 0
    m, mg, st = ctx.get_cells()
    assert (st == 0).all()
    assert np.abs(m - m0).max() <= 1e-12 * m0.max()


@pytest.mark.parametrize("periodic", [False, True])
def test_slab_clustered_seeds_flag_never_wrong(Context, periodic):
    """Clustered seeds on the slab path (two-level local grid, sweep of far-reaching cells): a cell whose
    security ball leaves the local grid is re-queued with a wider halo and, if it still does at the widest
    one, comes back UNCERTIFIED -- every other cell equals the single-context cell."""
    n = 60000
    xyz = oracles.clustered_seeds(n, 4321)
    m0, _, st0, _, _, _ = _single(Context, xyz, periodic, faces=False)
    ctx = Context(0)
    if periodic:
        ctx.set_domain_periodic(1.0)
    ctx.slab_set_points(xyz, 0, n)
    ctx.slab_compute_cells()
    m, _, st = ctx.get_cells()
    assert ((st & ~np.uint8(1 | 8)) == 0).all(), np.bincount(st)
    ok = (st == 0) & (st0 == 0)
    assert ok.sum() >= n - 30
    assert (np.abs(m - m0)[ok] <= 1e-12 * m0[ok]).all()
    if not periodic:
        assert (st == 0).all() and ctx.slab_stats()["retries"] == 0


def test_slab_two_ranks_vs_single_context():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tools", "slab_check.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "slab check ok" in out.stdout
