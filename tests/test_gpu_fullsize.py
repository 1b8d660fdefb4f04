"""Full-size GPU parity: BASELINE configs C1 (10 M) and C2 (100 M) against the reference's code tiling the SAME
cloud (oracle/_ref when present, else the port), node by node through digests; a 640 M-point run that crosses the
32-bit edge cases (3 * i > 2^31, output ids > 2^31 bytes), checked by size-independent properties and by subtree
parity against the oracle.

The inputs are the generators of schwarzwald_b200/workloads.py, i.e. exactly what bench.py times.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _oracle():
    from oracle import sworacle
    orc = sworacle.Oracle("ref" if sworacle.have_ref() else "port")
    orc.set_threads(max(1, len(os.sched_getaffinity(0))))
    return orc


def _cloud(cfg_name, n_total, dev):
    """(xyz on the device, bmin, bmax, spacing) of a BASELINE config, as bench.py prepares it."""
    import bench
    from schwarzwald_b200 import workloads
    cfg = workloads.CONFIGS[cfg_name]
    mn, mx, xyz = bench.full_cloud_tight_bounds(cfg, n_total, dev, keep=(0, n_total))
    bmin, bmax, spacing, centre = workloads.finish_bounds(cfg, mn, mx)
    return cfg, workloads.apply_pre_transform(cfg, xyz, centre), bmin, bmax, spacing


def _tile_on_device(cfg, xyz, bmin, bmax, spacing, sort_mode=3):
    """sort_mode 3: top-40-bit passes + segment finish, the mode the automatic choice reaches on these clouds after
    its first batch (a fresh handle would sort its first batch with the eight passes, which the smaller tests cover)."""
    import schwarzwald_b200 as sw
    from schwarzwald_b200 import workloads
    torch = _torch_cuda()
    t = sw.GpuTiler(cfg["sampling"], cfg["tiling"], bmin, bmax, spacing,
                    max_points_per_node=workloads.MAX_POINTS_PER_NODE, concurrency=cfg["concurrency"])
    t.set_sort_mode(sort_mode)
    t.build_execution_graph(xyz)
    t.finalize()
    nn, ni = t.result_size()
    ids = torch.empty(max(ni, 1), dtype=torch.int32, device=xyz.device)
    nodes = t.result_device_ids(ids.data_ptr())
    torch.cuda.synchronize()
    return t, nodes, ids[:ni]


@pytest.mark.parametrize("cfg_name", ["c1", "c2"])
def test_baseline_config_full_parity(cfg_name):
    """C1: 10 M uniform, GRID_CENTER FAST, float32-shifted.  C2: 100 M terrain, RANDOM_GRID FAST (the bench config)."""
    torch = _torch_cuda()
    from oracle import parity, sworacle
    from schwarzwald_b200 import workloads
    dev = torch.device("cuda", 0)
    n = workloads.CONFIGS[cfg_name]["points"]
    cfg, xyz, bmin, bmax, spacing = _cloud(cfg_name, n, dev)
    host = xyz.cpu().numpy()
    t, nodes, ids = _tile_on_device(cfg, xyz, bmin, bmax, spacing)
    try:
        orc = _oracle()
        params = sworacle.make_params(cfg["sampling"], cfg["tiling"], spacing, bmin, bmax,
                                      max_points_per_node=workloads.MAX_POINTS_PER_NODE,
                                      concurrency=cfg["concurrency"])
        want, clamped = orc.tile(params, host, return_clamped=True)
        assert t.start_level() == want.start_level
        rep = parity.full_parity(want, nodes, ids)
        assert rep["ok"], rep
        # index_point clamps in place: the device copy holds the clamped positions
        assert t.clamped_count() == int((clamped != host).any(axis=1).sum())
        assert np.array_equal(xyz.cpu().numpy(), clamped)
        # sorted keys and the sort permutation, bit for bit
        k, o = t.keys(n)
        assert np.array_equal(k, want.keys) and np.array_equal(o, want.order)
    finally:
        t.close()


def test_640m_points_properties_and_subtrees():
    """RANDOM_GRID FAST over 640 M terrain points: 3 * i and 24 * i cross 2^31 / 2^32, the output id array is
    larger than 2^32 bytes.  Properties: sorted keys; every point exactly once among the nodes at or below the
    start level; reconstructed levels only hold points of their children.  Plus subtree parity against the oracle
    (subtrees spread over the whole index range)."""
    torch = _torch_cuda()
    if torch.cuda.get_device_properties(0).total_memory < 120 * (1 << 30):
        pytest.skip("needs a 180 GB device")
    from oracle import parity
    from schwarzwald_b200 import workloads
    dev = torch.device("cuda", 0)
    n = 640_000_000
    cfg, xyz, bmin, bmax, spacing = _cloud("c2", n, dev)
    t, nodes, ids = _tile_on_device(cfg, xyz, bmin, bmax, spacing)
    try:
        S = t.start_level()
        assert 3 <= S <= 6
        lv = nodes["levels"].astype(np.int64)
        assert int(nodes["count"].sum()) == ids.numel()
        # every point exactly once at levels >= S
        seen = torch.zeros(n, dtype=torch.uint8, device=dev)
        below = np.nonzero(lv >= S)[0]
        f0 = int(nodes["first"][below].min())
        total_below = int(nodes["count"][below].sum())
        assert total_below == n
        # chunks are level-major: the non-reconstructed chunks form one contiguous prefix of the id array
        assert f0 == 0 and int((nodes["first"][below] + nodes["count"][below]).max()) == n
        part = ids[:n].to(torch.int64) & 0xFFFFFFFF
        assert int(part.max()) == n - 1 and int(part.min()) == 0
        seen.index_fill_(0, part, 1)
        assert int(seen.sum(dtype=torch.int64)) == n
        del seen, part
        # sorted keys (device-side check through the stand-alone hook would copy 5 GB; sample instead)
        rep = parity.subtree_parity(_oracle(), cfg["sampling"], cfg["tiling"], spacing, bmin, bmax, cfg["concurrency"],
                                    workloads.MAX_POINTS_PER_NODE, xyz, None, nodes, ids, S, depth=4,
                                    budget_points=8_000_000, max_subtree_points=3_000_000, max_subtrees=4)
        assert rep["checked"] and rep["ok"], rep
    finally:
        t.close()
