"""Host-side point orders that feed the multilevel drivers: the BRIO order and its levels
(compute_BRIO_order, mesh/mesh_reorder.cpp:1113-1155, 1287-1315: random shuffle, then recursively the
first `ratio` of the interval, the rest of every interval sorted along a Hilbert curve, `levels` = the end
of every interval) -- what the reference's sampling hands to OptimalTransportMap3d::optimize_levels
(src/examples/exploragram/compute_OTM/main.cpp:265-297).

Same interval arithmetic as the reference (so `levels` is identical for the same n, threshold, ratio);
inside an interval the points are sorted by their index on the 3-D Hilbert curve of a 2^16 lattice over the
bounding box (Skilling's transpose encoding) instead of the reference's recursive median splits: the same
role (consecutive points are close), not the same permutation.  The shuffle uses numpy's generator, not
GEO::random_shuffle.  On the device, b200vor_get_hilbert_order orders the bins of the grid the same way."""
import numpy as np


def hilbert_keys(xyz, bits=16, box=None):
    """Index of every point on the 3-D Hilbert curve of the 2^bits lattice over `box`
    ((lo, hi) arrays; default: the bounding box of the points).  Returns uint64 keys."""
    xyz = np.asarray(xyz, np.float64)
    lo = xyz.min(axis=0) if box is None else np.asarray(box[0], np.float64)
    hi = xyz.max(axis=0) if box is None else np.asarray(box[1], np.float64)
    ext = np.where(hi > lo, hi - lo, 1.0)
    q = np.clip(((xyz - lo) / ext * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    X = [q[:, 0].astype(np.uint64), q[:, 1].astype(np.uint64), q[:, 2].astype(np.uint64)]
    one = np.uint64(1)
    # axes -> transpose (J. Skilling, "Programming the Hilbert curve", 2004): inverse undo, then Gray encode
    Q = np.uint64(1 << (bits - 1))
    while Q > one:
        P = Q - one
        for i in range(3):
            hit = (X[i] & Q) != 0
            X[0] = np.where(hit, X[0] ^ P, X[0])                 # invert the low bits of x
            t = np.where(hit, np.uint64(0), (X[0] ^ X[i]) & P)   # else exchange the low bits of x and axis i
            X[0] = X[0] ^ t
            X[i] = X[i] ^ t
        Q = Q >> one
    X[1] ^= X[0]
    X[2] ^= X[1]
    t = np.zeros_like(X[0])
    Q = np.uint64(1 << (bits - 1))
    while Q > one:
        t = np.where((X[2] & Q) != 0, t ^ (Q - one), t)
        Q = Q >> one
    for i in range(3):
        X[i] ^= t
    # interleave the transpose: bit b of axis i is bit 3 b + (2 - i) of the key
    key = np.zeros_like(X[0])
    for b in range(bits):
        for i in range(3):
            key |= ((X[i] >> np.uint64(b)) & one) << np.uint64(3 * b + (2 - i))
    return key


def brio_levels(n, threshold=300, ratio=0.125):
    """The interval bounds of compute_BRIO_order: [0, e_1, ..., n] (mesh_reorder.cpp:1126-1154)."""
    bounds = []

    def rec(b, e):
        m = b
        if e - b > threshold:
            m = b + int(float(e - b) * ratio)
            rec(b, m)
        bounds.append((m, e))

    if n > 0:
        rec(0, n)
    return [0] + [e for _, e in bounds], bounds


def brio_order(xyz, threshold=300, ratio=0.125, seed=0):
    """Returns (order, levels): xyz[order] is in BRIO order, every prefix [0, levels[l]) a uniform random
    sub-sample whose last interval is sorted along the Hilbert curve."""
    xyz = np.asarray(xyz, np.float64)
    n = xyz.shape[0]
    order = np.random.default_rng(seed).permutation(n)
    levels, bounds = brio_levels(n, threshold, ratio)
    if n == 0:
        return order, [0]
    box = (xyz.min(axis=0), xyz.max(axis=0))
    for m, e in bounds:
        if e - m > 1:
            seg = order[m:e]
            keys = hilbert_keys(xyz[seg], box=box)
            order[m:e] = seg[np.argsort(keys, kind="stable")]
    return order, levels
