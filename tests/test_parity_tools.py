"""CPU tests of the full-size parity machinery itself (schwarzwald_b200/verify.py, oracle/parity.py).

The subtree method claims: for the grid strategies the part of EVERY node that lies inside a Morton-prefix subtree
only depends on the subtree's points (given the start level of the whole cloud); for MIN_DISTANCE FAST the same
holds for the nodes inside the subtree.  Here the claim is tested on the oracle alone: the whole cloud tiled by the
oracle plays the role of the GPU result.
"""
import numpy as np
import pytest
import torch

from oracle import parity, sworacle
from schwarzwald_b200 import synth, verify
from schwarzwald_b200.tiler import NODE_DTYPE, cubic_bounds, spacing_from_diagonal_fraction


def test_digest_host_and_torch_agree():
    rng = np.random.default_rng(1)
    counts = rng.integers(0, 50, size=2000)
    counts[5] = 0
    nodes = np.zeros(len(counts), NODE_DTYPE)
    nodes["first"] = np.concatenate([[0], np.cumsum(counts)[:-1]])
    nodes["count"] = counts
    nodes["levels"] = rng.integers(0, 6, len(counts))
    nodes["index"] = rng.integers(0, 1 << 40, len(counts))
    ids = rng.integers(0, 2 ** 32, size=int(counts.sum()), dtype=np.uint64).astype(np.uint32)
    a = verify.node_digests(nodes, ids)
    b = verify.node_digests_device(nodes, torch.from_numpy(ids.view(np.int32)))
    assert verify.compare_digests(a, b)[0]
    # order inside a node matters, the order of the node table does not
    k = int(np.argmax(counts))
    f = int(nodes["first"][k])
    swapped = ids.copy()
    swapped[f], swapped[f + 1] = ids[f + 1], ids[f]
    assert not verify.compare_digests(verify.node_digests(nodes, swapped), a)[0]
    perm = rng.permutation(len(nodes))
    assert verify.compare_digests(verify.node_digests(nodes[perm], ids), a)[0]
    assert verify.result_digest(a) == verify.result_digest(verify.node_digests(nodes[perm], ids))


def test_subtree_prefixes_match_the_oracle_keys(port_oracle):
    xyz = synth.generate("terrain", 200_000, 2, device="cpu", side_m=800.0)
    bmin, bmax = cubic_bounds(xyz.amin(0).numpy(), xyz.amax(0).numpy())
    xyz[:10] += 5000.0  # outliers: clamped by index_point
    keys, _ = port_oracle.index_points(xyz.numpy(), (bmin, bmax))
    for depth in (1, 3, 5):
        got = verify.subtree_prefixes_device(xyz, bmin, bmax, depth).numpy().astype(np.uint64)
        assert np.array_equal(got, keys >> np.uint64(3 * (21 - depth)))


@pytest.mark.parametrize("sampling,tiling", [("RANDOM_GRID", "FAST"), ("RANDOM_GRID", "ACCURATE"),
                                             ("GRID_CENTER", "ACCURATE"), ("JITTERED", "FAST"),
                                             ("MIN_DISTANCE", "FAST")])
def test_subtree_method_on_the_oracle_itself(port_oracle, sampling, tiling):
    n = 1_200_000 if not sampling.startswith("MIN") else 500_000
    xyz = synth.generate("urban", n, 3, device="cpu", side_m=600.0, height_m=60.0, n_primitives=300)
    bmin, bmax = cubic_bounds(xyz.amin(0).numpy(), xyz.amax(0).numpy())
    spacing = spacing_from_diagonal_fraction(bmin, bmax)
    params = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=2000, concurrency=4)
    whole = port_oracle.tile(params, xyz.numpy())
    rep = parity.subtree_parity(port_oracle, sampling, tiling, spacing, bmin, bmax, 4, 2000, xyz, None, whole.nodes,
                                whole.ids, whole.start_level, depth=3, budget_points=400_000,
                                max_subtree_points=200_000, max_subtrees=3)
    assert rep["checked"] and rep["ok"], rep
    assert rep["nodes"] > 10 and rep["ids"] > 1000
    # a single wrong id is noticed
    bad = whole.ids.copy()
    sub = rep["subtrees"][0]["prefix"]
    inside = np.nonzero((whole.nodes["levels"] >= 3) &
                        ((whole.nodes["index"] >> (3 * (whole.nodes["levels"].astype(np.uint64) - np.uint64(3)))) == sub) &
                        (whole.nodes["count"] > 1))[0]
    f = int(whole.nodes["first"][inside[0]])
    bad[f], bad[f + 1] = bad[f + 1], bad[f]
    rep2 = parity.subtree_parity(port_oracle, sampling, tiling, spacing, bmin, bmax, 4, 2000, xyz, None, whole.nodes,
                                 bad, whole.start_level, depth=3, budget_points=400_000,
                                 max_subtree_points=200_000, max_subtrees=3)
    assert not rep2["ok"]


def test_full_parity_report(port_oracle):
    xyz = synth.generate("uniform", 100_000, 1, device="cpu", side_m=100.0).numpy()
    bmin, bmax = cubic_bounds(xyz.min(0), xyz.max(0))
    spacing = spacing_from_diagonal_fraction(bmin, bmax)
    params = sworacle.make_params("GRID_CENTER", "FAST", spacing, bmin, bmax, max_points_per_node=1000, concurrency=2)
    want = port_oracle.tile(params, xyz)
    rep = parity.full_parity(want, want.nodes, want.ids)
    assert rep["ok"] and rep["nodes"] == len(want.nodes)
    rep = parity.full_parity(want, want.nodes, torch.from_numpy(want.ids.view(np.int32)))
    assert rep["ok"]


def test_min_spacing_check(port_oracle):
    xyz = synth.generate("terrain", 300_000, 2, device="cpu", side_m=500.0).numpy()
    bmin, bmax = cubic_bounds(xyz.min(0), xyz.max(0))
    spacing = spacing_from_diagonal_fraction(bmin, bmax)
    params = sworacle.make_params("MIN_DISTANCE", "ACCURATE", spacing, bmin, bmax, max_points_per_node=2000,
                                  concurrency=2)
    res = port_oracle.tile(params, xyz)
    rep = parity.min_spacing_check(lambda ids: xyz[ids], res.nodes, res.ids, spacing)
    assert rep["ok"] and len(rep["nodes"]) >= 3, rep
    # RANDOM_GRID does not keep the spacing: the check must notice
    params = sworacle.make_params("RANDOM_GRID", "ACCURATE", spacing, bmin, bmax, max_points_per_node=2000,
                                  concurrency=2)
    res = port_oracle.tile(params, xyz)
    rep = parity.min_spacing_check(lambda ids: xyz[ids], res.nodes, res.ids, spacing)
    assert not rep["ok"]


def test_min_spacing_helpers_use_the_reference_threshold():
    """too_close_pairs / merged_min_spacing_check (bench.py's MIN_DISTANCE legs): squared distance strictly below
    (float)(spacing_f * spacing_f) of the node's level, SparseGrid.cpp:11-14 / GridCell.cpp:41-58."""
    from oracle import parity
    s = np.float32(0.5)
    thr = float(np.float32(s * s))
    d_equal = np.sqrt(thr)  # exactly at the threshold: not a conflict (strict <)
    pts = np.array([[0, 0, 0], [d_equal, 0, 0], [10, 0, 0], [10, d_equal * 0.999, 0]], np.float64)
    dx = pts[1] - pts[0]
    expect = int((dx @ dx) < thr) + 1
    assert parity.too_close_pairs(pts, 0, s) == expect
    assert parity.too_close_pairs(pts, 1, s) == 0  # one level deeper: half the spacing
    assert parity.too_close_pairs(pts[:1], 0, s) == 0
    # a node split over two ranks: the conflict only shows on the merged node
    parts = [{(0, 0): pts[[2]]}, {(0, 0): pts[[3]], (1, 3): pts[[0]]}]
    rep = parity.merged_min_spacing_check(parts, s)
    assert rep["checked"] and not rep["ok"] and rep["too_close_pairs"] == 1 and rep["n_nodes"] == 2
    assert parity.merged_min_spacing_check([{}, {}], s)["checked"] is False


def test_workload_configs_match_baseline_json():
    """schwarzwald_b200/workloads.py names the five BASELINE.json configs with their sizes and strategies."""
    import json
    import os
    from schwarzwald_b200 import workloads
    base = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "BASELINE.json")))
    sizes = {"c1": 10_000_000, "c2": 100_000_000, "c3": 500_000_000, "c4": 1_000_000_000, "c5": 2_000_000_000}
    words = {"c1": ("GRID_CENTER", "FAST"), "c2": ("RANDOM_GRID", "FAST"), "c3": ("JITTERED", "FAST"),
             "c4": ("MIN_DISTANCE", "ACCURATE"), "c5": ("GRID_CENTER",)}
    for k, (name, text) in enumerate(zip(sorted(sizes), base["configs"])):
        cfg = workloads.CONFIGS[name]
        assert cfg["points"] == sizes[name]
        for w in words[name]:
            assert w in text and w in (cfg["sampling"], cfg["tiling"])
    assert workloads.default_config(1) == "c2" and workloads.default_config(8) == "c3"
