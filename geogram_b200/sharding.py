"""Host-side sharding of the cell path over the GPUs of one node.

Cells are independent given the seeds, so rank r computes the cells of the user-index range
[first, last) and every rank holds all seeds.  The only exchange per Lloyd iteration is the
all-gather of the updated rows (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""


def shard_range(n_total, rank, world):
    """Contiguous, balanced user-index range of `rank`: sizes differ by at most one and the
    ranges partition [0, n_total)."""
    base, rem = divmod(int(n_total), int(world))
    first = rank * base + min(rank, rem)
    last = first + base + (1 if rank < rem else 0)
    return first, last


def shard_sizes(n_total, world):
    return [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]


def slab_order(full):
    """Permutation that sorts the seeds by x: with `shard_range`, rank r then owns the r-th
    x-quantile SLAB of the domain (equal seed counts, cuts at the quantiles -- SURVEY 8(e)), so the
    cells a rank computes are spatially coherent and share their neighbourhoods in L1/L2.
    Deterministic (stable sort), hence identical on every rank."""
    import torch
    return torch.argsort(full[:, 0], stable=True)


def allgather_rows(dist, full, mine):
    """full[n_total, c] <- concatenation over ranks of mine[(last-first), c].

    Equal shards use all_gather_into_tensor (one NCCL ring/NVLS op); ragged shards are padded to the largest one."""
    world = dist.get_world_size()
    n_total = full.shape[0]
    sizes = shard_sizes(n_total, world)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(full, mine.contiguous())
        return full
    # ragged: pad every shard to the largest one, gather, drop the padding
    import torch
    smax = max(sizes)
    pad = torch.zeros((smax,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    pad[: mine.shape[0]] = mine
    scratch = torch.empty((world * smax,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    dist.all_gather_into_tensor(scratch, pad)
    at = 0
    for r, sz in enumerate(sizes):
        full[at:at + sz] = scratch[r * smax: r * smax + sz]
        at += sz
    return full


class ShardedLloyd(object):
    """CVT Lloyd relaxation of one seed set over the GPUs of one node (one process per GPU).

    Rank r owns the rows [first, last) of the seed array (`shard_range`).  Every iteration:
    each rank computes the cells of its rows against ALL seeds (b200vor_set_query_range +
    b200vor_lloyd), then the updated rows are all-gathered (NCCL over NVLink) and the grid is
    rebuilt on every rank.  This is the multi-GPU form of
    CentroidalVoronoiTesselation::Lloyd_iterations (voronoi/CVT.cpp:133-169); the reference has no
    distributed counterpart (single address space).

    `dist` is torch.distributed (already initialised) or None for a single process."""

    def __init__(self, ctx, n_total, dist=None, device=None, stream=None):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.n_total = int(n_total)
        self.first, self.last = shard_range(self.n_total, self.rank, self.world)
        self.device = device
        self.full = torch.empty((self.n_total, 3), dtype=torch.float64, device=device)
        # staging of the updated seeds (only the own rows are used) when sharded
        self.out = torch.empty((self.n_total, 3), dtype=torch.float64, device=device) if self.world > 1 else None
        # gloo / CPU tensors (host-logic tests with a stand-in context): no stream, no events
        self.on_gpu = self.full.is_cuda
        self.stream = None
        if self.on_gpu:
            self.stream = stream if stream is not None else torch.cuda.current_stream(device)
            ctx.set_stream(self.stream.cuda_stream)
        if self.world > 1:
            ctx.set_query_range(self.first, self.last)
        self.phase = {"ms_grid": 0.0, "ms_knn": 0.0, "ms_clip": 0.0}
        self._comm = []
        self._stale = False     # sharded: the context still holds the seeds of the previous step

    # ---- seeds in ----
    def set_seeds_device(self, full, slabs=False):
        """full: (n_total,3) float64 device tensor holding ALL seeds (identical on every rank).
        slabs=True first re-orders the rows into x-quantile slabs (`slab_order`); the permutation
        is kept in `self.order` (row k of the job is row order[k] of the caller)."""
        self.order = None
        with self._on_stream():
            if slabs:
                self.order = slab_order(full)
                self.full.copy_(full[self.order])
            else:
                self.full.copy_(full)
        self.ctx.set_points_device(self.n_total, self.full.data_ptr())
        self._stale = False
        self._take("ms_grid")

    def set_shard_host(self, host_rows):
        """host_rows: this rank's rows as a (last-first,3) float64 host tensor (pinned for speed).
        H2D of the own rows only; the other ranks' rows arrive over NVLink."""
        mine = self.full[self.first:self.last]
        with self._on_stream():
            mine.copy_(host_rows, non_blocking=True)
            if self.world > 1:
                self._gather(mine)
        self.ctx.set_points_device(self.n_total, self.full.data_ptr())
        self._stale = False
        self._take("ms_grid")

    def _refresh(self):
        """Grid of the all-gathered seeds, built when the next step needs it (a caller that
        uploads new seeds every step never pays for the rebuild after the step)."""
        if self._stale:
            self.ctx.set_points_device(self.n_total, self.full.data_ptr())
            self._stale = False
            self._take("ms_grid")

    # ---- one Lloyd iteration ----
    def step(self, k0=0, mode=0):
        if self.world == 1:
            self.ctx.lloyd_device(1, None, k0=k0, mode=mode)    # seeds stay resident, updated in place
            self._take("ms_grid", "ms_knn", "ms_clip")
            return
        self._refresh()
        self.ctx.lloyd_device(1, self.out.data_ptr(), k0=k0, mode=mode)
        self._take("ms_knn", "ms_clip")
        with self._on_stream():
            self._gather(self.out[self.first:self.last])
        self._stale = True          # the grid of the new seeds is built lazily (_refresh)

    def get_shard_host(self, host_rows):
        """Own rows of the current seeds -> host tensor (D2H of the own rows only)."""
        if self.world == 1:
            self.ctx.lloyd_device(0, self.full.data_ptr())
        with self._on_stream():
            host_rows.copy_(self.full[self.first:self.last], non_blocking=True)
        if self.on_gpu:
            self.stream.synchronize()
        return host_rows

    def comm_ms(self):
        """Device time of the all-gathers since the last call (CUDA events on the stream)."""
        if not self.on_gpu:
            return 0.0
        self.torch.cuda.synchronize(self.device)
        ms = sum(a.elapsed_time(b) for a, b in self._comm)
        self._comm = []
        return ms

    # ---- internals ----
    def _on_stream(self):
        """Every torch op of the job (copies, NCCL) is issued on the stream the context is bound
        to, whatever torch's current stream is at call time."""
        if self.on_gpu:
            return self.torch.cuda.stream(self.stream)
        import contextlib
        return contextlib.nullcontext()

    def _gather(self, mine):
        if not self.on_gpu:
            allgather_rows(self.dist, self.full, mine)
            return
        a = self.torch.cuda.Event(enable_timing=True)
        b = self.torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        allgather_rows(self.dist, self.full, mine)
        b.record(self.stream)
        self._comm.append((a, b))

    def _take(self, *keys):
        s = self.ctx.stats()
        for k in keys:
            self.phase[k] += s[k]


class SlabLloyd(object):
    """CVT Lloyd relaxation over the GPUs of one node with x-SLABS AND HALOS (one process per GPU):
    the driver of b200vor_slab_* (csrc/slab.cu).  Rank r keeps only its own rows [first, last) of the
    x-sorted seed array on the device; every iteration it exchanges the seeds within a halo of the
    neighbouring slabs (ncclSend / ncclRecv inside the library), bins own + ghost seeds on a
    slab-local grid and updates its own seeds.  No rank ever holds, bins or receives all seeds.
    Same interface as ShardedLloyd (the replicated variant above).

    `dist` (torch.distributed, initialised) is used ONCE, to hand the NCCL id of the library's own
    communicator to every rank."""

    def __init__(self, ctx, n_total, dist=None, device=None, stream=None, halo_bins=None):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.n_total = int(n_total)
        self.first, self.last = shard_range(self.n_total, self.rank, self.world)
        self.device = device
        self.stream = stream if stream is not None else torch.cuda.current_stream(device)
        ctx.set_stream(self.stream.cuda_stream)
        if self.world > 1:
            from .capi import Context
            uid = [Context.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.comm_init(uid[0], self.rank, self.world)
        if halo_bins:
            ctx.slab_set_halo_bins(halo_bins)
        self.mine = torch.empty((self.last - self.first, 3), dtype=torch.float64, device=device)
        self.order = None
        self.phase = {"ms_grid": 0.0, "ms_knn": 0.0, "ms_clip": 0.0}
        self._comm_ms = 0.0
        self.retries = 0
        self.last_stats = None
        self.steps = []          # per step: device ms of exchange / pack / grid / clip, ghosts, retries

    def set_seeds_device(self, full, slabs=True):
        """full: (n_total,3) float64 device tensor of ALL seeds, identical on every rank (set-up
        only): sorted by x once, this rank keeps its rows."""
        with self.torch.cuda.stream(self.stream):
            self.order = slab_order(full) if slabs else None
            rows = full[self.order[self.first:self.last]] if slabs else full[self.first:self.last]
            self.mine.copy_(rows)
        self.stream.synchronize()
        self.ctx.slab_set_points_device(self.last - self.first, self.first, self.n_total, self.mine.data_ptr())

    def set_shard_host(self, host_rows):
        with self.torch.cuda.stream(self.stream):
            self.mine.copy_(host_rows, non_blocking=True)
        self.ctx.slab_set_points_device(self.last - self.first, self.first, self.n_total, self.mine.data_ptr())

    def step(self, k0=0, mode=0):
        self.ctx.slab_lloyd_device(1, None)
        s = self.ctx.slab_stats()
        self.last_stats = s
        self.phase["ms_grid"] += s["ms_grid"]
        self.phase["ms_clip"] += s["ms_clip"]
        self._comm_ms += s["ms_exchange"] + s["ms_pack"]
        self.retries += s["retries"]
        self.steps.append({k: (round(s[k], 3) if isinstance(s[k], float) else s[k])
                           for k in ("ms_exchange", "ms_pack", "ms_grid", "ms_clip", "n_local", "retries", "reallocs")})

    def get_shard_host(self, host_rows):
        self.ctx.slab_lloyd_device(0, self.mine.data_ptr())
        with self.torch.cuda.stream(self.stream):
            host_rows.copy_(self.mine, non_blocking=True)
        self.stream.synchronize()
        return host_rows

    def comm_ms(self):
        ms, self._comm_ms = self._comm_ms, 0.0
        return ms


# ---- host statement of the halo selection of csrc/slab.cu (XPred / YZPred), numpy: what a rank must
#      hold to compute its own cells.  Used by the CPU tests (gloo) and as the specification of the
#      device kernels; the product path never calls it. ----
def home_image(x, center, period):
    """image of the x-period that brings x next to `center` (slab.cu: home_image)"""
    import numpy as np
    return -np.rint((x - center) / period).astype(np.int64)


def circular_center(x, period):
    import numpy as np
    a = 2.0 * np.pi * x / period
    c = np.arctan2(np.sin(a).sum(), np.cos(a).sum()) / (2.0 * np.pi)
    return (c + 1.0 if c < 0.0 else c) * period


def slab_bounds(x_own, period=None):
    """(lo, hi, center) of a rank's own seeds; periodic: an arc around the circular mean"""
    if period is None:
        return float(x_own.min()), float(x_own.max()), 0.0
    c = circular_center(x_own, period)
    xu = x_own + home_image(x_own, c, period) * period
    return float(xu.min()), float(xu.max()), float(c)


def halo_send_list(xyz_own, first, rank, bounds, H, period=None):
    """For every peer p: (rows of (x', y, z), global ids, x-images) of the own seeds (and their
    x-images) within [lo_p - H, hi_p + H] -- own-index order inside each (peer, image) bucket."""
    import numpy as np
    x = xyz_own[:, 0]
    s0 = home_image(x, bounds[rank][2], period) if period is not None else np.zeros(len(x), np.int64)
    out = []
    for p, (lo, hi, _) in enumerate(bounds):
        rows, gids, imgs = [], [], []
        for s in (-1, 0, 1):
            if s != 0 and period is None:
                continue
            xs = x + (s * period if period is not None else 0.0)
            sel = (xs >= lo - H) & (xs <= hi + H)
            if p == rank:
                sel &= s0 != s
            idx = np.nonzero(sel)[0]
            r = xyz_own[idx].copy()
            r[:, 0] = xs[idx]
            rows.append(r)
            gids.append(first + idx)
            imgs.append(np.full(len(idx), s, np.int64))
        out.append((np.concatenate(rows), np.concatenate(gids), np.concatenate(imgs)))
    return out


def halo_local_set(xyz_own, first, rank, bounds, received, H, period=None):
    """Local seed array of a rank: own rows (at their home image), received ghosts in peer order,
    then (periodic) the y / z images within H of a face.  Returns (xyz, global id, (tx, ty, tz))."""
    import numpy as np
    own = xyz_own.copy()
    tx0 = np.zeros(len(own), np.int64)
    if period is not None:
        tx0 = home_image(own[:, 0], bounds[rank][2], period)
        own[:, 0] += tx0 * period
    xyz = np.concatenate([own] + [r[0] for r in received])
    gid = np.concatenate([first + np.arange(len(own))] + [r[1] for r in received])
    tx = np.concatenate([tx0] + [r[2] for r in received])
    t = np.stack([tx, np.zeros_like(tx), np.zeros_like(tx)], axis=1)
    if period is None:
        return xyz, gid, t
    add_xyz, add_gid, add_t = [], [], []
    for tz in (-1, 0, 1):
        for ty in (-1, 0, 1):
            if ty == 0 and tz == 0:
                continue
            y = xyz[:, 1] + ty * period
            z = xyz[:, 2] + tz * period
            sel = (y >= -H) & (y <= period + H) & (z >= -H) & (z <= period + H)
            if ty == -1:
                sel &= xyz[:, 1] >= period - H
            if ty == 1:
                sel &= xyz[:, 1] <= H
            if tz == -1:
                sel &= xyz[:, 2] >= period - H
            if tz == 1:
                sel &= xyz[:, 2] <= H
            idx = np.nonzero(sel)[0]
            r = xyz[idx].copy()
            r[:, 1] = y[idx]
            r[:, 2] = z[idx]
            add_xyz.append(r)
            add_gid.append(gid[idx])
            tt = t[idx].copy()
            tt[:, 1] = ty
            tt[:, 2] = tz
            add_t.append(tt)
    return (np.concatenate([xyz] + add_xyz), np.concatenate([gid] + add_gid), np.concatenate([t] + add_t))
