"""Pins the oracle restatement against the reference's OWN hot-path code.

Two routes:
  * live: oracle/_ref/libswref.so = the reference's translation units compiled verbatim from
    /root/reference (only where that tree or a prebuilt library is present);
  * committed golden fixtures under tests/golden/ that were generated from that library by
    tests/golden/make_golden.py (always available, also on the GPU box).
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SAMPLINGS = ["RANDOM_GRID", "GRID_CENTER", "JITTERED", "MIN_DISTANCE", "MIN_DISTANCE_FAST"]


def cloud(seed, n, scale, offset):
    rng = np.random.default_rng(seed)
    return np.round(rng.random((n, 3)) * np.array(scale) + np.array(offset), 3)


@pytest.mark.parametrize("tiling", ["ACCURATE", "FAST"])
@pytest.mark.parametrize("sampling", SAMPLINGS)
def test_port_equals_reference_whole_batch(port_oracle, ref_oracle, sampling, tiling):
    import schwarzwald_b200 as sw
    from oracle import sworacle
    xyz = cloud(21, 120_000, [900.0, 700.0, 80.0], [4000.0, -250.0, 10.0])
    xyz[:7] += 3000.0  # outliers get clamped
    bmin, bmax = sw.cubic_bounds(xyz[7:].min(0), xyz[7:].max(0))
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    p = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=1500, concurrency=3)
    a, ca = port_oracle.tile(p, xyz, return_clamped=True)
    b, cb = ref_oracle.tile(p, xyz, return_clamped=True)
    assert np.array_equal(ca, cb)
    assert np.array_equal(a.keys, b.keys) and np.array_equal(a.order, b.order)
    assert a.start_level == b.start_level
    ta, ia = a.canonical()
    tb, ib = b.canonical()
    assert np.array_equal(ta, tb) and np.array_equal(ia, ib)


@pytest.mark.parametrize("sampling", SAMPLINGS)
@pytest.mark.parametrize("node_level", [-1, 0, 2])
def test_port_equals_reference_sample_points(port_oracle, ref_oracle, sampling, node_level):
    """sample_points() on one node, both behaviours."""
    rng = np.random.default_rng(5 + node_level)
    bounds = (np.array([-3.0, 10.0, 100.0]), np.array([61.0, 74.0, 164.0]))
    # points inside the first octant chain so that they share the node prefix
    pts = bounds[0] + rng.random((4000, 3)) * (64.0 / 2 ** (node_level + 1))
    keys, _ = port_oracle.index_points(pts, bounds)
    order = np.argsort(keys, kind="stable").astype(np.uint32)
    for behaviour in (0, 1):
        ra = port_oracle.sample_points(sampling, pts, keys[order], order, 0, node_level, bounds, 1.7,
                                       behaviour=behaviour, max_points_per_node=5000)
        rb = ref_oracle.sample_points(sampling, pts, keys[order], order, 0, node_level, bounds, 1.7,
                                      behaviour=behaviour, max_points_per_node=5000)
        assert ra[0] == rb[0]
        assert np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2])
        if behaviour == 0:
            assert ra[0] == len(pts)  # TakeAllWhenCountBelowMaxPoints


def test_port_equals_reference_primitives(port_oracle, ref_oracle):
    rng = np.random.default_rng(8)
    bounds = (np.array([12.5, -7.25, 3.0]), np.array([112.5, 92.75, 103.0]))
    pts = rng.random((20000, 3)) * 130.0 - 10.0
    ka, ca = port_oracle.index_points(pts, bounds)
    kb, cb = ref_oracle.index_points(pts, bounds)
    assert np.array_equal(ka, kb) and np.array_equal(ca, cb)
    for key in ka[:50]:
        for depth in (1, 5, 13, 21):
            a = port_oracle.bounds_from_morton_index(key, depth, bounds)
            b = ref_oracle.bounds_from_morton_index(key, depth, bounds)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for v in rng.integers(0, 1 << 21, 100):
        assert port_oracle.expand_bits_by_3(int(v)) == ref_oracle.expand_bits_by_3(int(v))
    for sampling in SAMPLINGS:
        for spacing in (0.3, 0.69, 1.0, 7.5, 40.0):
            for level in range(-1, 14):
                assert port_oracle.required_morton_index_depth(sampling, level, bounds, spacing) == \
                    ref_oracle.required_morton_index_depth(sampling, level, bounds, spacing)
    sk = np.sort(ka)
    for level in (0, 1, 3):
        assert np.array_equal(port_oracle.partition_child_octants(sk, level),
                              ref_oracle.partition_child_octants(sk, level))


def test_port_equals_reference_jitter_errors(port_oracle, ref_oracle):
    """Both raise the reference's JITTERED exception for grids below 16 cells (Sampling.h:632-635).
    The second exception ("node too small", Sampling.h:641-653) is only exercised on the port: in
    this image the reference's own message formatting for it (boost::format shim + to_string of a
    DynamicMortonIndex) crashes nondeterministically, which is outside the compute path."""
    from oracle import sworacle
    rng = np.random.default_rng(3)
    pts = rng.random((3000, 3))
    bounds = (np.zeros(3), np.ones(3))
    keys, _ = port_oracle.index_points(pts, bounds)
    order = np.argsort(keys, kind="stable").astype(np.uint32)
    for orc in (port_oracle, ref_oracle):
        with pytest.raises(RuntimeError, match="code 4"):
            orc.sample_points("JITTERED", pts, keys[order], order, 0, -1, bounds, 0.1, behaviour=1)
    with pytest.raises(RuntimeError, match="code 5"):
        port_oracle.sample_points("JITTERED", pts, keys[order], order, 0, 17, bounds, 0.01, behaviour=1)


def load_golden():
    path = os.path.join(HERE, "golden", "tiler_golden.json")
    return json.load(open(path))


@pytest.mark.parametrize("case", range(20))
def test_port_equals_committed_reference_golden(port_oracle, case):
    """Fixtures generated from oracle/_ref (the reference's own code) by tests/golden/make_golden.py."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    from oracle import sworacle
    g = load_golden()["cases"][case]
    xyz, bmin, bmax, spacing = make_golden.case_input(g["cloud"])
    p = sworacle.make_params(g["sampling"], g["tiling"], spacing, bmin, bmax, max_points_per_node=g["max_points"],
                             concurrency=g["concurrency"])
    res = port_oracle.tile(p, xyz)
    digest = make_golden.digest(res)
    assert digest == g["digest"], (g["sampling"], g["tiling"], g["cloud"])


# ---------------------------------------------------------------------------------------------------
# SURVEY section 8 f1 (groundwork): several batches through TilingAlgorithmV1 with cached points
# ---------------------------------------------------------------------------------------------------
def _batch_cloud():
    import schwarzwald_b200 as sw
    xyz = cloud(33, 90_000, [400.0, 300.0, 60.0], [1000.0, -20.0, 5.0])
    xyz[:5] -= 700.0  # outliers of the first batch get clamped
    bmin, bmax = sw.cubic_bounds(xyz[5:].min(0), xyz[5:].max(0))
    return xyz, bmin, bmax, sw.spacing_from_diagonal_fraction(bmin, bmax)


@pytest.mark.parametrize("sampling", SAMPLINGS)
def test_multi_batch_port_equals_reference(port_oracle, ref_oracle, sampling):
    """read_pnts_from_disk re-keying (the reference's own calculate_morton_index relative to the node bounds in
    the ref build), merge with the incoming points, AlwaysAdhereToMinSpacing on revisits."""
    from oracle import sworacle
    xyz, bmin, bmax, spacing = _batch_cloud()
    p = sworacle.make_params(sampling, "ACCURATE", spacing, bmin, bmax, max_points_per_node=800, concurrency=2)
    sizes = [20_000, 45_000, 25_000]
    a, ca = port_oracle.tile_batches(p, xyz, sizes, return_clamped=True)
    b, cb = ref_oracle.tile_batches(p, xyz, sizes, return_clamped=True)
    assert np.array_equal(ca, cb)
    ta, ia = a.canonical()
    tb, ib = b.canonical()
    assert np.array_equal(ta, tb) and np.array_equal(ia, ib)
    seen = np.bincount(a.ids, minlength=len(xyz))
    assert (seen == 1).all(), "every point is stored in exactly one node"


@pytest.mark.parametrize("sampling", ["RANDOM_GRID", "JITTERED", "MIN_DISTANCE"])
def test_one_batch_through_the_batch_path_equals_single_batch(port_oracle, sampling):
    from oracle import sworacle
    xyz, bmin, bmax, spacing = _batch_cloud()
    p = sworacle.make_params(sampling, "ACCURATE", spacing, bmin, bmax, max_points_per_node=800, concurrency=2)
    one = port_oracle.tile(p, xyz)
    same = port_oracle.tile_batches(p, xyz, [len(xyz)])
    t1, i1 = one.canonical()
    t2, i2 = same.canonical()
    assert np.array_equal(t1[:, :3], t2[:, :3]) and np.array_equal(i1, i2)


def test_multi_batch_revisited_nodes_are_resampled(port_oracle):
    """A node that already holds points is sampled with AlwaysAdhereToMinSpacing on the next visit: the union of
    two small batches that would each be stored whole (count <= max_points_per_node) is thinned out."""
    from oracle import sworacle
    xyz, bmin, bmax, spacing = _batch_cloud()
    xyz = xyz[5:6005]
    p = sworacle.make_params("GRID_CENTER", "ACCURATE", spacing, bmin, bmax, max_points_per_node=4000, concurrency=2)
    first = port_oracle.tile_batches(p, xyz[:3000], [3000])
    assert len(first.nodes) == 1 and int(first.nodes["count"][0]) == 3000  # taken whole
    both = port_oracle.tile_batches(p, xyz, [3000, 3000])
    root = both.nodes[(both.nodes["levels"] == 0)]
    assert len(root) == 1 and int(root["count"][0]) < 6000 and len(both.nodes) > 1
    assert (np.bincount(both.ids, minlength=len(xyz)) == 1).all()


@pytest.mark.parametrize("sampling", SAMPLINGS)
def test_multi_batch_fast_port_equals_reference(port_oracle, ref_oracle, sampling):
    """TilingAlgorithmV3 over several batches: start level from the first batch, later batches cut at it and
    merged with the stored start nodes, reconstruction over every start node at the end."""
    from oracle import sworacle
    xyz, bmin, bmax, spacing = _batch_cloud()
    p = sworacle.make_params(sampling, "FAST", spacing, bmin, bmax, max_points_per_node=800, concurrency=2)
    sizes = [40_000, 30_000, 20_000]
    a = port_oracle.tile_batches(p, xyz, sizes)
    b = ref_oracle.tile_batches(p, xyz, sizes)
    assert a.start_level == b.start_level >= 3
    ta, ia = a.canonical()
    tb, ib = b.canonical()
    assert np.array_equal(ta, tb) and np.array_equal(ia, ib)
    below = a.nodes["levels"] >= a.start_level  # reconstructed upper levels hold copies
    ids = np.concatenate([a.ids[int(n["first"]): int(n["first"]) + int(n["count"])] for n in a.nodes[below]])
    assert (np.bincount(ids, minlength=len(xyz)) == 1).all()
    one = port_oracle.tile(p, xyz)
    same = port_oracle.tile_batches(p, xyz, [len(xyz)])
    t1, i1 = one.canonical()
    t2, i2 = same.canonical()
    assert np.array_equal(t1[:, :3], t2[:, :3]) and np.array_equal(i1, i2)


def _batches_golden():
    return json.load(open(os.path.join(HERE, "golden", "batches_golden.json")))["cases"]


@pytest.mark.parametrize("case", range(10))
def test_multi_batch_port_equals_committed_reference_golden(port_oracle, case):
    """Fixtures generated from oracle/_ref by tests/golden/make_golden_batches.py (available without the
    reference tree)."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_batches as mgb
    g = _batches_golden()[case]
    res = mgb.run(port_oracle, g["sampling"], g["tiling"])
    assert (len(res.nodes), len(res.ids), res.start_level) == (g["nodes"], g["ids"], g["start_level"])
    assert mgb.digest(res) == g["digest"], (g["sampling"], g["tiling"])
