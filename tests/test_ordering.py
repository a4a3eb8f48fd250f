"""BRIO order and levels (geogram_b200/ordering.py; mesh/mesh_reorder.cpp:1113-1155, 1287-1315)."""
import numpy as np

from geogram_b200.ordering import brio_levels, brio_order, hilbert_keys


def test_hilbert_keys_walk_the_lattice_one_step_at_a_time():
    """All 8^b lattice points sorted by key: consecutive points are lattice neighbours, keys are 0..8^b-1
    (the defining property of the Hilbert curve)."""
    b = 4
    g = np.arange(1 << b)
    P = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float64)
    box = (np.zeros(3), np.full(3, float(1 << b)))
    k = hilbert_keys(P + 0.5, bits=b, box=box)
    assert np.array_equal(np.sort(k), np.arange(8 ** b, dtype=np.uint64))
    walk = P[np.argsort(k)]
    assert (np.abs(np.diff(walk, axis=0)).sum(axis=1) == 1).all()


def _reference_levels(n, threshold, ratio):
    # independent restatement of the recursion as a loop: the nested prefixes from the outside in
    ends = [n]
    while ends[-1] > threshold:
        ends.append(int(float(ends[-1]) * ratio))
    return [0] + ends[::-1]


def test_brio_levels_follow_the_reference_interval_arithmetic():
    for n, th, r in [(10000000, 300, 0.125), (4000, 300, 0.125), (299, 300, 0.125), (301, 300, 0.125),
                     (123457, 64, 0.25), (1, 300, 0.125)]:
        levels, bounds = brio_levels(n, th, r)
        assert levels == _reference_levels(n, th, r)
        assert bounds[0][0] == 0 and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:])) and bounds[-1][1] == n
    assert brio_levels(10000000)[0] == [0, 38, 305, 2441, 19531, 156250, 1250000, 10000000]
    assert brio_levels(0)[0] == [0]


def test_brio_order_is_a_permutation_with_uniform_prefixes_and_local_intervals():
    rng = np.random.default_rng(9)
    n = 20000
    xyz = rng.random((n, 3))
    xyz[:, 0] = xyz[:, 0] ** 2           # non-uniform density
    order, levels = brio_order(xyz, seed=3)
    assert np.array_equal(np.sort(order), np.arange(n)) and levels[0] == 0 and levels[-1] == n
    # every prefix is a uniform sub-sample: its x-quantiles are those of the whole set
    for e in levels[2:-1]:
        qs = np.quantile(xyz[order[:e], 0], [0.25, 0.5, 0.75])
        qa = np.quantile(xyz[:, 0], [0.25, 0.5, 0.75])
        assert np.abs(qs - qa).max() < 4.0 / np.sqrt(e)
    # inside the last interval consecutive points are close (Hilbert), unlike a random order
    m = levels[-2]
    seg = xyz[order[m:]]
    step = np.linalg.norm(np.diff(seg, axis=0), axis=1).mean()
    rnd = np.linalg.norm(np.diff(xyz[rng.permutation(n)[: n - m]], axis=0), axis=1).mean()
    assert step < 0.15 * rnd
    # same seed, same order
    assert np.array_equal(order, brio_order(xyz, seed=3)[0])


def test_brio_levels_equal_the_unmodified_reference():
    """tests/golden/brio_reference.npz: compute_BRIO_order of the unmodified reference (make_golden.py brio)."""
    import os
    from conftest import GOLDEN
    G = np.load(os.path.join(GOLDEN, "brio_reference.npz"))
    for n, th, r in G["cases"]:
        ref = G["levels_%d_%d_%g" % (int(n), int(th), r)].tolist()
        assert brio_levels(int(n), int(th), float(r))[0] == ref
    # same kind of order: the last interval of the reference's order and of ours are equally local
    xyz, oref = G["xyz_20000"], G["order_20000"].astype(np.int64)
    order, levels = brio_order(xyz, seed=1)
    m = levels[-2]

    def step(o):
        return np.linalg.norm(np.diff(xyz[o[m:]], axis=0), axis=1).mean()
    assert step(order) < 1.25 * step(oref)
    assert set(order.tolist()) == set(oref.tolist())
