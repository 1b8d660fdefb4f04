"""GPU multi-batch tiling against the CPU oracle and the committed reference digests.

swgpu_set_multi_batch: every build_execution_graph() is one batch of TilingAlgorithmV1 / V3 with cached points
(tile_node, TilingAlgorithms.cpp:351-492; read_pnts_from_disk, :50-109; merge_node_data_*, Node.cpp:3-34; FAST
later iterations, :1362-1453; reconstruct over the store, :1717-1784).  The checker is the oracle's
tile_batches (oracle/orchestrator.h run_accurate_batches / run_fast_batches), itself pinned against the verbatim
reference build by tests/test_oracle_vs_reference.py and tests/golden/batches_golden.json.
"""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
SAMPLINGS = ["RANDOM_GRID", "GRID_CENTER", "JITTERED", "MIN_DISTANCE", "MIN_DISTANCE_FAST"]


def _torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def gpu_tile_batches(sampling, tiling, xyz, sizes, bmin, bmax, spacing, max_pts, conc, max_depth=100, device_input=False):
    """Feeds xyz batch by batch through the C ABI; returns (TileResult with global ids, clamped positions)."""
    import schwarzwald_b200 as sw
    host = np.array(xyz, dtype=np.float64, order="C", copy=True)
    with sw.GpuTiler(sampling, tiling, bmin, bmax, spacing, max_points_per_node=max_pts, concurrency=conc,
                     max_depth=max_depth) as t:
        t.set_multi_batch(True)
        lo = 0
        for n in sizes:
            part = host[lo:lo + n]
            if device_input:
                torch = _torch_cuda()
                dev = torch.from_numpy(part).cuda()
                t.build_execution_graph(dev)
                torch.cuda.synchronize()
                host[lo:lo + n] = dev.cpu().numpy()
            else:
                t.build_execution_graph(part)  # clamps in place, like index_point
            lo += n
        assert lo == len(host)
        t.finalize()
        res = t.result()
    return res, host


def _golden_mod():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_batches as mgb
    return mgb


def _golden_cases():
    return json.load(open(os.path.join(HERE, "golden", "batches_golden.json")))["cases"]


@pytest.mark.parametrize("case", range(10))
def test_multi_batch_matches_reference_golden_and_port(port_oracle, case):
    """All five strategies x {ACCURATE, FAST}, three batches, five clamped outliers in the first one: digests
    generated from the verbatim reference build, and the port live."""
    _torch_cuda()
    from oracle import sworacle
    mgb = _golden_mod()
    g = _golden_cases()[case]
    xyz, bmin, bmax, spacing = mgb.case_input()
    sizes = mgb.SPLITS[g["tiling"]]
    got, clamped = gpu_tile_batches(g["sampling"], g["tiling"], xyz, sizes, bmin, bmax, spacing, 800, 2)
    p = sworacle.make_params(g["sampling"], g["tiling"], spacing, bmin, bmax, max_points_per_node=800, concurrency=2)
    want, want_clamped = port_oracle.tile_batches(p, xyz, sizes, return_clamped=True)
    assert np.array_equal(clamped, want_clamped)
    assert got.start_level == want.start_level
    wt, wi = want.canonical()
    gt, gi = got.canonical()
    assert np.array_equal(wt, gt), "node table (levels, index, count, flags) differs"
    assert np.array_equal(wi, gi), "node contents differ"
    assert (len(got.nodes), len(got.ids), got.start_level) == (g["nodes"], g["ids"], g["start_level"])
    assert mgb.digest(got) == g["digest"]


@pytest.mark.parametrize("tiling", ["ACCURATE", "FAST"])
@pytest.mark.parametrize("sampling", ["RANDOM_GRID", "GRID_CENTER", "JITTERED", "MIN_DISTANCE"])
def test_multi_batch_terrain_many_batches(port_oracle, sampling, tiling):
    """600 k terrain points in 7 uneven batches, device-resident input."""
    _torch_cuda()
    import schwarzwald_b200 as sw
    from oracle import sworacle
    from schwarzwald_b200 import synth
    xyz = synth.generate("terrain", 600_000, 11, device="cpu", side_m=2000.0).numpy()
    bmin, bmax = sw.cubic_bounds(xyz.min(0), xyz.max(0))
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    sizes = [150_000, 50_000, 100_000, 7, 99_993, 120_000, 80_000]
    if tiling == "FAST":
        sizes[3], sizes[4] = 8, 99_992  # parallel::scatter needs >= concurrency points per batch
    got, _ = gpu_tile_batches(sampling, tiling, xyz, sizes, bmin, bmax, spacing, 5000, 8, device_input=True)
    p = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=5000, concurrency=8)
    want = port_oracle.tile_batches(p, xyz, sizes)
    assert got.start_level == want.start_level
    wt, wi = want.canonical()
    gt, gi = got.canonical()
    assert np.array_equal(wt, gt) and np.array_equal(wi, gi)
    lv = got.nodes["levels"]
    keep = lv >= max(got.start_level, 0)
    ids = np.concatenate([got.ids[int(n["first"]): int(n["first"]) + int(n["count"])] for n in got.nodes[keep]])
    assert (np.bincount(ids, minlength=len(xyz)) == 1).all(), "every point is stored exactly once"


@pytest.mark.parametrize("sampling", SAMPLINGS)
def test_multi_batch_terminal_nodes(port_oracle, sampling):
    """max_depth = 2 with a coarse spacing (diagonal / 60): nodes at level 2 are terminal; on a revisit they
    concatenate incoming and stored points unsorted (merge_node_data_unsorted, Node.cpp:22-34)."""
    _torch_cuda()
    import schwarzwald_b200 as sw
    from oracle import sworacle
    from schwarzwald_b200 import synth
    xyz = synth.generate("uniform", 100_000, 4, device="cpu", side_m=50.0).numpy()
    bmin, bmax = sw.cubic_bounds(xyz.min(0), xyz.max(0))
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax, 60.0)
    sizes = [40_000, 35_000, 25_000]
    got, _ = gpu_tile_batches(sampling, "ACCURATE", xyz, sizes, bmin, bmax, spacing, 100, 2, max_depth=2)
    p = sworacle.make_params(sampling, "ACCURATE", spacing, bmin, bmax, max_points_per_node=100, concurrency=2,
                             max_depth=2)
    want = port_oracle.tile_batches(p, xyz, sizes)
    wt, wi = want.canonical()
    gt, gi = got.canonical()
    assert (wt[:, 3] & 2).any(), "the case must reach terminal nodes"
    assert np.array_equal(wt, gt) and np.array_equal(wi, gi)


@pytest.mark.parametrize("tiling", ["ACCURATE", "FAST"])
def test_one_batch_through_the_store_equals_single_batch(tiling):
    """The multi-batch path with one batch and the single-batch pipeline give the same nodes."""
    _torch_cuda()
    import schwarzwald_b200 as sw
    mgb = _golden_mod()
    xyz, bmin, bmax, spacing = mgb.case_input()
    got, _ = gpu_tile_batches("JITTERED", tiling, xyz, [len(xyz)], bmin, bmax, spacing, 800, 2)
    with sw.GpuTiler("JITTERED", tiling, bmin, bmax, spacing, max_points_per_node=800, concurrency=2) as t:
        one = t.tile(xyz.copy())
    t1, i1 = one.canonical()
    t2, i2 = got.canonical()
    assert np.array_equal(t1, t2) and np.array_equal(i1, i2)


def test_multi_batch_mode_resets_and_refuses_sharding():
    _torch_cuda()
    import schwarzwald_b200 as sw
    mgb = _golden_mod()
    xyz, bmin, bmax, spacing = mgb.case_input()
    with sw.GpuTiler("RANDOM_GRID", "ACCURATE", bmin, bmax, spacing, max_points_per_node=800, concurrency=2) as t:
        t.set_multi_batch(True)
        t.build_execution_graph(xyz[:20_000].copy())
        n1 = t.result_size()
        t.set_multi_batch(True)  # switching empties the store
        assert t.result_size() == (0, 0)
        t.build_execution_graph(xyz[:20_000].copy())
        assert t.result_size() == n1
        t.set_multi_batch(False)
        r = t.tile(xyz[:20_000].copy())
        assert int(r.nodes["count"].sum()) == 20_000


@pytest.mark.parametrize("sampling,tiling", [("RANDOM_GRID", "FAST"), ("JITTERED", "ACCURATE")])
def test_multi_batch_writer_payloads(port_oracle, sampling, tiling):
    """PNTS / LAS position payloads of the final node store (PNTSWriter.cpp:326-342, LASPersistence.h:119-131,160-163)
    over the positions of ALL batches, against the oracle's payload functions on the same nodes."""
    _torch_cuda()
    import schwarzwald_b200 as sw
    mgb = _golden_mod()
    xyz, bmin, bmax, spacing = mgb.case_input()
    sizes = mgb.SPLITS[tiling]
    host = np.array(xyz, dtype=np.float64, order="C", copy=True)
    with sw.GpuTiler(sampling, tiling, bmin, bmax, spacing, max_points_per_node=800, concurrency=2) as t:
        t.set_multi_batch(True)
        lo = 0
        for n in sizes:
            t.build_execution_graph(host[lo:lo + n])
            lo += n
        t.finalize()
        res = t.result()
        pnts = t.payload_pnts()
        las, headers = t.payload_las()
    assert np.array_equal(pnts, port_oracle.payload_pnts(host, res.ids))
    w_las, w_headers = port_oracle.payload_las(host, res.ids, res.nodes, (bmin, bmax))
    assert np.array_equal(las, w_las)
    for f in ("offset", "scale", "max"):
        assert np.array_equal(headers[f], w_headers[f])


def test_multi_batch_duplicates_reach_the_re_root_depth():
    """The same points arriving batch after batch can never be separated by a sampling grid: a revisited leaf is
    always sampled (TilingAlgorithms.cpp:272-275), keeps one copy and hands the others to a new child, which takes
    them whole - one level deeper per batch, until the depth where the reference re-roots (:444-483).  Default
    policy: SW_ERR_DEEP_REROOT, like the oracle.  swgpu_set_deep_node_policy(1): those nodes store their points
    whole, flagged TERMINAL | DEEP, every point exactly once."""
    _torch_cuda()
    import schwarzwald_b200 as sw
    from schwarzwald_b200 import tiler as swt
    mgb = _golden_mod()
    xyz, bmin, bmax, spacing = mgb.case_input()
    xyz = xyz[5:30_005].copy()
    copies = np.concatenate([xyz[:200]] * 3)
    sizes = [30_000] + [600] * 16
    cloud = np.concatenate([xyz] + [copies] * 16)
    for sampling in ("RANDOM_GRID", "JITTERED"):
        with sw.GpuTiler(sampling, "ACCURATE", bmin, bmax, spacing, max_points_per_node=800, concurrency=2) as t:
            t.set_multi_batch(True)
            with pytest.raises(sw.SwgpuError) as e:
                lo = 0
                for n in sizes:
                    t.build_execution_graph(cloud[lo:lo + n].copy())
                    lo += n
            assert e.value.code == 6
        with sw.GpuTiler(sampling, "ACCURATE", bmin, bmax, spacing, max_points_per_node=800, concurrency=2) as t:
            t.set_multi_batch(True)
            t.set_deep_node_policy(True)
            lo = 0
            for n in sizes:
                t.build_execution_graph(cloud[lo:lo + n].copy())
                lo += n
            t.finalize()
            res = t.result()
        assert (np.bincount(res.ids.astype(np.int64), minlength=len(cloud)) == 1).all()
        deep = res.nodes[(res.nodes["flags"] & swt.NODE_DEEP) != 0]
        assert len(deep) > 0 and (deep["flags"] & swt.NODE_TERMINAL).all() and (deep["levels"] == 15).all()
        assert int(deep["count"].sum()) < 49 * 200
