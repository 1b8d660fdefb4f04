"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the
hot path.  Each test cites the reference test it restates (paths relative to
/root/reference/schwarzwald/test).  The reference tests use MortonIndex<N> with small N; the same
geometry is expressed here in the 21-level key space (levels are counted from the root, so the top
levels of a 21-level key are the N-level key)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def octant_at_level(key, level):
    return (int(key) >> (3 * (20 - level))) & 7


def position_from_octant_indices(indices, bmin, bmax):
    """helper of TestOctreeIndexing.cpp:20-42"""
    mn = np.array(bmin, np.float64)
    mx = np.array(bmax, np.float64)
    for o in indices:
        for bit, axis in ((1, 2), (2, 1), (4, 0)):
            if o & bit:
                mn[axis] += (mx[axis] - mn[axis]) / 2
            else:
                mx[axis] -= (mx[axis] - mn[axis]) / 2
    return mn + (mx - mn) / 2


def test_expand_and_contract_bits(port_oracle):
    """stuff.h:207-234 — 21 input bits, every third output bit."""
    rng = np.random.default_rng(0)
    for v in [0, 1, 2, 0x1FFFFF, 0x155555, 0xAAAAA] + [int(x) for x in rng.integers(0, 1 << 21, 200)]:
        e = port_oracle.expand_bits_by_3(v)
        want = 0
        for b in range(21):
            want |= ((v >> b) & 1) << (3 * b)
        assert e == want
        assert port_oracle.contract_bits_by_3(e) == v
        assert port_oracle.contract_bits_by_3(e | (e << 1) | (e << 2)) == v


def test_first_level_octants(port_oracle):
    """TestOctreeIndexing.cpp:72-98: octants 0,1,2,4 in the unit cube."""
    pts = np.array([[0.25, 0.25, 0.25], [0.25, 0.25, 0.75], [0.25, 0.75, 0.25], [0.75, 0.25, 0.25]])
    keys, _ = port_oracle.index_points(pts, ([0, 0, 0], [1, 1, 1]))
    assert [octant_at_level(k, 0) for k in keys] == [0, 1, 2, 4]


def test_twenty_level_key(port_oracle):
    """TestOctreeIndexing.cpp:100-126 and :584-600."""
    octants = [5, 3, 7, 4, 0, 1, 6, 4, 3, 5, 3, 6, 7, 3, 2, 1, 4, 0, 2, 5]
    bounds = ([0, 0, 0], [1 << 20] * 3)
    p = position_from_octant_indices(octants, *bounds)
    keys, _ = port_oracle.index_points(p[None, :], bounds)
    assert [octant_at_level(keys[0], l) for l in range(20)] == octants


def test_position_from_octants_helper():
    """TestOctreeIndexing.cpp:44-69."""
    b = ([0, 0, 0], [8, 8, 8])
    assert position_from_octant_indices([0, 0], *b).tolist() == [1, 1, 1]
    assert position_from_octant_indices([3, 0], *b).tolist() == [1, 5, 5]
    assert position_from_octant_indices([0, 5], *b).tolist() == [3, 1, 3]


def test_index_point_clamps_outliers(port_oracle):
    """index_point writes the clamped coordinates back; a coordinate equal to max maps to 2^21-1
    (OctreeAlgorithms.h:74-79,156-170)."""
    pts = np.array([[2.0, -1.0, 0.5], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0]])
    keys, clamped = port_oracle.index_points(pts, ([0, 0, 0], [1, 1, 1]))
    assert clamped.tolist() == [[1.0, 0.0, 0.5], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0]]
    assert int(keys[1]) == (1 << 63) - 1 and int(keys[2]) == 0
    full = port_oracle.expand_bits_by_3((1 << 21) - 1)
    assert int(keys[0]) == (full << 2) | port_oracle.expand_bits_by_3(1 << 20)


def test_random_grid_known_answer(port_oracle):
    """TestOctreeIndexing.cpp:169-252: 32^3 lattice, spacing 32, node level 0, max 16 points ->
    exactly the 8 lattice points (0.5|16.5)^3, in Morton order."""
    side = 32
    g = np.arange(side) + 0.5
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    bounds = ([0, 0, 0], [side] * 3)
    keys, _ = port_oracle.index_points(pts, bounds)
    order = np.argsort(keys, kind="stable")
    n_sel, _, ids = port_oracle.sample_points("RANDOM_GRID", pts, keys[order], order.astype(np.uint32), 0, 0, bounds,
                                              float(side), max_points_per_node=16)
    assert n_sel == 8
    expected = [[0.5, 0.5, 0.5], [0.5, 0.5, 16.5], [0.5, 16.5, 0.5], [0.5, 16.5, 16.5], [16.5, 0.5, 0.5],
                [16.5, 0.5, 16.5], [16.5, 16.5, 0.5], [16.5, 16.5, 16.5]]
    assert pts[ids[:8]].tolist() == expected


@pytest.mark.parametrize("sampling", ["RANDOM_GRID", "GRID_CENTER", "JITTERED", "MIN_DISTANCE"])
def test_sampling_is_stable(port_oracle, sampling):
    """TestOctreeIndexing.cpp:254-336: selected and remainder both stay Morton-sorted."""
    rng = np.random.default_rng(1)
    pts = rng.random((5000, 3)) * 64
    bounds = ([0, 0, 0], [64] * 3)
    keys, _ = port_oracle.index_points(pts, bounds)
    order = np.argsort(keys, kind="stable")
    n_sel, k_out, _ = port_oracle.sample_points(sampling, pts, keys[order], order.astype(np.uint32), 0, -1, bounds, 2.0,
                                                max_points_per_node=16)
    assert 0 < n_sel < len(pts)
    assert (np.diff(k_out[:n_sel].astype(np.int64)) >= 0).all()
    assert (np.diff(k_out[n_sel:].astype(np.int64)) >= 0).all()


def test_partition_root_level(port_oracle):
    """TestOctreeIndexing.cpp:338-392: one point per octant of [0,4]^3."""
    pts = np.array([[1, 1, 1], [1, 1, 3], [1, 3, 1], [1, 3, 3], [3, 1, 1], [3, 1, 3], [3, 3, 1], [3, 3, 3]], float)
    keys, _ = port_oracle.index_points(pts, ([0, 0, 0], [4, 4, 4]))
    assert port_oracle.partition_child_octants(keys, 0).tolist() == list(range(9))


def test_partition_level_four(port_oracle):
    """TestOctreeIndexing.cpp:394-459: [0,2) (2,2) (2,2) [2,3) (3,3) [3,5) [5,6) (6,6)."""
    octs = [[3, 4, 5, 2, 0], [3, 4, 5, 2, 0], [3, 4, 5, 2, 3], [3, 4, 5, 2, 5], [3, 4, 5, 2, 5], [3, 4, 5, 2, 6]]
    bounds = ([0, 0, 0], [32] * 3)
    pts = np.array([position_from_octant_indices(o, *bounds) for o in octs])
    keys, _ = port_oracle.index_points(pts, bounds)
    assert port_oracle.partition_child_octants(keys, 4).tolist() == [0, 2, 2, 2, 3, 3, 5, 6, 6]


def test_octant_bounds(port_oracle):
    """TestOctreeIndexing.cpp:461-492."""
    expected = {0: ([0, 0, 0], [2, 2, 2]), 1: ([0, 0, 2], [2, 2, 4]), 2: ([0, 2, 0], [2, 4, 2]), 3: ([0, 2, 2], [2, 4, 4]),
                4: ([2, 0, 0], [4, 2, 2]), 5: ([2, 0, 2], [4, 2, 4]), 6: ([2, 2, 0], [4, 4, 2]), 7: ([2, 2, 2], [4, 4, 4])}
    for o, (emin, emax) in expected.items():
        mn, mx = port_oracle.octant_bounds(o, ([0, 0, 0], [4, 4, 4]))
        assert mn.tolist() == emin and mx.tolist() == emax


def test_points_inside_child_bounds(port_oracle):
    """TestOctreeIndexing.cpp:494-554."""
    rng = np.random.default_rng(2)
    pts = rng.integers(1024, 2049, (1024, 3)).astype(np.float64)
    bounds = ([1024] * 3, [2048] * 3)
    keys, _ = port_oracle.index_points(pts, bounds)
    order = np.argsort(keys, kind="stable")
    cuts = port_oracle.partition_child_octants(keys[order], 0)
    for o in range(8):
        mn, mx = port_oracle.octant_bounds(o, bounds)
        sub = pts[order[int(cuts[o]):int(cuts[o + 1])]]
        assert ((sub >= mn) & (sub <= mx)).all()


def test_bounds_from_morton_index(port_oracle):
    """TestOctreeIndexing.cpp:556-582."""
    mn, mx = port_oracle.bounds_from_morton_index(0, 1, ([0, 0, 0], [2, 2, 2]))
    assert mn.tolist() == [0, 0, 0] and mx.tolist() == [1, 1, 1]
    key = (1 << 60) | (4 << 57) | (5 << 54)
    mn, mx = port_oracle.bounds_from_morton_index(key, 3, ([0, 0, 0], [8, 8, 8]))
    assert mn.tolist() == [3, 0, 5] and mx.tolist() == [4, 1, 6]


def test_node_names():
    """TestMortonIndex.cpp:81-140 (Potree 'r' prefix) and TestOctreeNodeIndex.cpp:434-447
    (Entwine "13-410-7041-4059" <-> {2,3,1,3,7,7,1,0,5,5,0,5,3})."""
    import schwarzwald_b200 as sw
    octs = [2, 3, 1, 3, 7, 7, 1, 0, 5, 5, 0, 5, 3]
    idx = 0
    for o in octs:
        idx = (idx << 3) | o
    assert sw.node_name(idx, len(octs), "entwine") == "13-410-7041-4059"
    assert sw.node_name(idx, len(octs)) == "r" + "".join(map(str, octs))
    assert sw.node_name(0, 0) == "r"
    assert sw.node_name((7 << 9) | (6 << 6) | (0 << 3) | 3, 4) == "r7603"  # TestMortonIndex.cpp:37-52


def test_jitter_tables_match_fixture():
    """Both generated copies of PERMUTATIONS_16/32/64 (Sampling.h:14-138) equal the committed fixture."""
    fixture = json.load(open(os.path.join(HERE, "golden", "jitter_tables.json")))
    import re
    for path in ("oracle/jitter_tables.inc", "schwarzwald_b200/csrc/jitter_tables.cuh"):
        src = open(os.path.join(HERE, "..", path)).read()
        for size in ("16", "32", "64"):
            body = re.search(r"PERMUTATIONS_%s\[16\]\[%s\] = \{(.*?)\};" % (size, size), src, re.S).group(1)
            rows = [[int(v) for v in r.split(",")] for r in re.findall(r"\{([0-9, ]+)\}", body)]
            assert rows == fixture[size]
            for r in rows:
                assert sorted(r) == list(range(1, int(size) + 1))


def test_required_morton_index_depth_defaults(port_oracle):
    """Default spacing = diagonal / 250: grid strategies sample 7 levels below the node, JITTERED
    uses a 128-cell grid (7 levels), MIN_DISTANCE needs no deeper keys (Sampling.cpp:29-62)."""
    import schwarzwald_b200 as sw
    bounds = (np.zeros(3), np.full(3, 1000.0))
    spacing = sw.spacing_from_diagonal_fraction(*bounds)
    for level in range(-1, 12):
        assert port_oracle.required_morton_index_depth("RANDOM_GRID", level, bounds, spacing) == level + 7
        assert port_oracle.required_morton_index_depth("GRID_CENTER", level, bounds, spacing) == level + 7
        assert port_oracle.required_morton_index_depth("JITTERED", level, bounds, spacing) == level + 7
        assert port_oracle.required_morton_index_depth("MIN_DISTANCE", level, bounds, spacing) == level
