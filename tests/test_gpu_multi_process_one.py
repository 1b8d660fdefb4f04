"""swgpu_multi_*: several GPUs driven from ONE host process (one thread per GPU, peer access instead of NCCL) —
the path the reference-side adapter uses, since the reference is a single process (process/Tiler.cpp:189-198).
On a box with one GPU the ranks share it (`devices` names it several times); with more GPUs they spread out."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devices(k):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n = torch.cuda.device_count()
    return [r % n for r in range(k)]


def _cloud(kind, n, seed, **kw):
    import schwarzwald_b200 as sw
    from schwarzwald_b200 import synth
    xyz = synth.generate(kind, n, seed, device="cpu", **kw).numpy()
    bmin, bmax = sw.cubic_bounds(xyz.min(0), xyz.max(0))
    return xyz, bmin, bmax, sw.spacing_from_diagonal_fraction(bmin, bmax)


def _same(want, got):
    assert got.start_level == want.start_level
    wt, wi = want.canonical()
    gt, gi = got.canonical()
    assert np.array_equal(wt[:, :3], gt[:, :3]), "node table differs"
    assert np.array_equal(wt[:, 3] & 7, gt[:, 3] & 7), "node flags differ"
    assert np.array_equal(wi, gi), "node contents differ"


@pytest.mark.parametrize("ranks", [2, 3])
@pytest.mark.parametrize("tiling", ["ACCURATE", "FAST"])
@pytest.mark.parametrize("sampling", ["RANDOM_GRID", "GRID_CENTER", "JITTERED"])
def test_one_process_many_gpus_bit_exact(port_oracle, sampling, tiling, ranks):
    import schwarzwald_b200 as sw
    from oracle import sworacle
    xyz, bmin, bmax, spacing = _cloud("terrain", 500_000, 2, side_m=1500.0)
    xyz[:7] -= 5000.0  # outliers: clamped on the GPU that holds the slice, written back into the caller's buffer
    params = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=3000, concurrency=4)
    want, clamped = port_oracle.tile(params, xyz, return_clamped=True)
    host = xyz.copy()
    with sw.MultiGpuTiler(sampling, tiling, bmin, bmax, spacing, _devices(ranks), max_points_per_node=3000,
                          concurrency=4) as t:
        got = t.tile(host)
        info = t.info()
    assert np.array_equal(host, clamped)
    assert info["clamped"] == 7 and int(info["shard_points"].sum()) == len(xyz)
    _same(want, got)


def test_one_process_attributes_and_reuse(port_oracle):
    """Two batches through the same handle (buffers are reused), attributes travelling with the points: every
    GPU's node-major attribute payload equals the attributes of its points' ids."""
    import schwarzwald_b200 as sw
    from oracle import sworacle
    rng = np.random.default_rng(3)
    with_t = None
    for n, seed in ((300_000, 4), (200_000, 5)):
        xyz, bmin, bmax, spacing = _cloud("urban", n, seed, side_m=600.0)
        attrs = rng.integers(0, 256, size=(n, 16), dtype=np.uint8)
        params = sworacle.make_params("GRID_CENTER", "FAST", spacing, bmin, bmax, max_points_per_node=2500, concurrency=4)
        want = port_oracle.tile(params, xyz)
        t = sw.MultiGpuTiler("GRID_CENTER", "FAST", bmin, bmax, spacing, _devices(3), max_points_per_node=2500,
                             concurrency=4)
        try:
            for _ in range(2):
                got = t.tile(xyz.copy(), attrs)
                _same(want, got)
            seen = 0
            for r in range(3):
                part, a = t.rank_result_with_attributes(r)
                assert np.array_equal(a, attrs[part.ids.astype(np.int64)])
                seen += len(part.ids)
            assert seen == len(got.ids)
        finally:
            t.close()


def test_one_process_min_distance_invariant(port_oracle):
    """MIN_DISTANCE over 3 ranks of one process: shard faces resolved through the in-process all-gather."""
    import schwarzwald_b200 as sw
    from oracle import parity, sworacle
    xyz, bmin, bmax, spacing = _cloud("terrain", 600_000, 2, side_m=1500.0)
    with sw.MultiGpuTiler("MIN_DISTANCE", "ACCURATE", bmin, bmax, spacing, _devices(3), max_points_per_node=3000,
                          concurrency=4) as t:
        got = t.tile(xyz.copy())
    assert (np.bincount(got.ids.astype(np.int64), minlength=len(xyz)) == 1).all()
    bad = 0
    for nd in got.nodes:
        if nd["flags"] & 3 or nd["count"] < 2:
            continue
        ids = got.ids[int(nd["first"]): int(nd["first"]) + int(nd["count"])].astype(np.int64)
        bad += parity.too_close_pairs(xyz[ids], int(nd["levels"]), spacing)
    assert bad == 0
    params = sworacle.make_params("MIN_DISTANCE", "ACCURATE", spacing, bmin, bmax, max_points_per_node=3000, concurrency=4)
    want = port_oracle.tile(params, xyz)
    root_ref = int(want.nodes[want.nodes["levels"] == 0]["count"][0])
    root_got = int(got.nodes[got.nodes["levels"] == 0]["count"][0])
    assert abs(root_got - root_ref) <= 0.01 * root_ref + 8


def test_one_process_errors_do_not_hang():
    """A reference exception on one rank (JITTERED: grid smaller than 16x16, Sampling.h:632-635) comes back as its
    error code; the other ranks are released from their barriers."""
    import schwarzwald_b200 as sw
    xyz, bmin, bmax, _ = _cloud("uniform", 100_000, 6, side_m=50.0)
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax, 20.0)
    with sw.MultiGpuTiler("JITTERED", "ACCURATE", bmin, bmax, spacing, _devices(2), max_points_per_node=100,
                          concurrency=2) as t:
        with pytest.raises(sw.SwgpuError) as e:
            t.tile(xyz.copy())
        assert e.value.code in (4, 1)
    with sw.MultiGpuTiler("RANDOM_GRID", "FAST", bmin, bmax, sw.spacing_from_diagonal_fraction(bmin, bmax), _devices(2),
                          concurrency=8) as t:
        with pytest.raises(sw.SwgpuError) as e:
            t.tile(xyz[:5].copy())
        assert e.value.code == 9
