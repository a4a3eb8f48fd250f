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
