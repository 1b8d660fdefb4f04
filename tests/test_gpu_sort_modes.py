"""K2 in its two forms: eight LSD passes, or passes over the top digits + the segment finish kernel
(swgpu_set_sort_mode).  Every mode must produce the order std::sort + the id tie rule produce
(TilingAlgorithms.cpp:600-604; SURVEY.md section 8a "S"): bit-exact keys and permutation.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIN_TILE = 4096   # kernels_index_sort.cu
FIN_LIMIT = 256


def _torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _sort_on_gpu(keys, mode):
    torch = _torch_cuda()
    import schwarzwald_b200 as sw
    n = len(keys)
    dev = torch.from_numpy(keys.view(np.int64).copy()).cuda()
    order = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
    one = np.array([0.0, 0.0, 0.0]), np.array([1.0, 1.0, 1.0])
    with sw.GpuTiler("RANDOM_GRID", "ACCURATE", one[0], one[1], 0.1) as t:
        t.set_sort_mode(mode)
        t.sort_keys_device(dev.data_ptr(), n, order.data_ptr())
        torch.cuda.synchronize()
        stats = t.stats()
    return dev.cpu().numpy().view(np.uint64), order.cpu().numpy().view(np.uint32)[:n], stats


def _check(keys, mode, fallback=None):
    got_keys, got_order, stats = _sort_on_gpu(keys, mode)
    perm = np.argsort(keys, kind="stable")
    assert np.array_equal(got_keys, keys[perm]), "sorted keys differ (mode %d)" % mode
    assert np.array_equal(got_order, perm.astype(np.uint32)), "permutation differs (mode %d)" % mode
    if mode:
        assert stats["sort_first_bit"] == 8 * mode
        if fallback is not None:  # 0 none, 1 eight more LSD passes, 2 long runs sorted one by one
            assert stats["sort_fallback"] == int(fallback), stats
            assert stats["sort_passes"] == (8 - mode) + (8 if int(fallback) == 1 else 0), stats
    else:
        assert stats["sort_passes"] == 8 and stats["sort_first_bit"] == 0, stats
    return stats


def _segments(n, seg_len, low_bits, rng, offset=0, lo_range=None, shuffle=True):
    """n keys whose top bits form runs of exactly seg_len elements (the first one shortened by `offset`)."""
    hi = (np.arange(n, dtype=np.uint64) + np.uint64(offset)) // np.uint64(seg_len)
    lo = rng.integers(0, lo_range or (1 << low_bits), n, dtype=np.uint64)
    keys = (hi * np.uint64(977) + np.uint64(5)) << np.uint64(low_bits) | lo
    assert int(keys.max()) < (1 << 63)
    if shuffle:
        keys = keys[rng.permutation(n)]
    return keys


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_random_keys_every_mode(mode):
    rng = np.random.default_rng(11 + mode)
    keys = rng.integers(0, 1 << 63, 1_000_003, dtype=np.uint64)
    _check(keys, mode, fallback=False)


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("seg_len", [1, 2, 3, 7, 64, 255, 256])
def test_runs_up_to_the_limit_are_finished_in_place(mode, seg_len):
    rng = np.random.default_rng(seg_len * 10 + mode)
    n = 5 * FIN_TILE + 1234
    for offset in (0, 1, seg_len // 2, seg_len - 1):  # run boundaries against the tile boundaries
        keys = _segments(n, seg_len, 8 * mode, rng, offset=offset)
        st = _check(keys, mode, fallback=False)
        if seg_len == 1:
            assert st["sort_moved"] == 0


@pytest.mark.parametrize("mode", [1, 3])
def test_ties_keep_the_original_index_order(mode):
    rng = np.random.default_rng(5)
    n = 3 * FIN_TILE + 77
    keys = _segments(n, 200, 8 * mode, rng, lo_range=4)  # about 50 equal keys per value
    _check(keys, mode, fallback=False)


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("seg_len", [257, 300, 4096, 10_000])
def test_long_unsorted_runs_fall_back_to_the_full_sort(mode, seg_len):
    rng = np.random.default_rng(seg_len + mode)
    n = 6 * FIN_TILE + 5
    for offset in (0, 17):
        keys = _segments(n, seg_len, 8 * mode, rng, offset=offset)
        _check(keys, mode, fallback=1)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_a_few_long_runs_are_sorted_one_by_one(mode):
    """Dense spots: runs of 257 .. 10 000 elements (several tiles) between short runs hold less than 1/8 of the keys."""
    rng = np.random.default_rng(40 + mode)
    low = 8 * mode
    parts, hi = [], 1

    def run(length, lo_range=None, ordered=False):
        nonlocal hi
        lo = rng.integers(0, lo_range or (1 << low), length, dtype=np.uint64)
        if ordered:
            lo.sort()
        parts.append((np.uint64(hi) << np.uint64(low)) | lo)
        hi += 3

    run(300)                      # a long run at the very start
    for length in (257, 1000, 5000, 10_000, 700):
        for _ in range(40_000 // 7):
            run(7)
        run(length)
    run(4000, ordered=True)       # long, but in order already
    run(600, lo_range=3)          # long with many ties
    for _ in range(3000):
        run(1)
    run(258)                      # a long run at the very end
    keys = np.concatenate(parts)
    # the input order: runs stay together only after the top-digit passes, so shuffle everything
    keys = keys[rng.permutation(len(keys))]
    st = _check(keys, mode, fallback=2)
    assert st["sort_passes"] == 8 - mode


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_one_inversion_deep_inside_a_long_run(mode):
    """The run covers several tiles; only two neighbouring elements are out of order."""
    n = 4 * FIN_TILE
    for where in (1, FIN_LIMIT - 1, FIN_LIMIT, FIN_TILE - 1, FIN_TILE, FIN_TILE + 1, FIN_TILE + FIN_LIMIT - 1,
                  FIN_TILE + FIN_LIMIT, 2 * FIN_TILE + 100, n - 1):
        lo = np.arange(n, dtype=np.uint64) // np.uint64(65)  # non-decreasing, < 253
        keys = (np.uint64(42) << np.uint64(8 * mode)) | lo
        a, b = where - 1, where
        keys[a] = (np.uint64(42) << np.uint64(8 * mode)) | (lo[b] + np.uint64(1))  # the only descent: a -> b
        d = np.diff(keys.astype(np.int64))
        assert (d < 0).sum() == 1 and d[a] < 0
        _check(keys, mode, fallback=1)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_long_runs_already_in_order_need_no_fallback(mode):
    n = 3 * FIN_TILE + 9
    same = np.full(n, (123 << (8 * mode)) | 7, dtype=np.uint64)          # identical points
    _check(same, mode, fallback=False)
    rising = (np.uint64(9) << np.uint64(8 * mode)) | (np.arange(n, dtype=np.uint64) // np.uint64(64))
    _check(rising, mode, fallback=False)
    rng = np.random.default_rng(1)
    mixed = np.concatenate([same[:5000], _segments(20_000, 9, 8 * mode, rng) + (np.uint64(1) << np.uint64(40))])
    _check(mixed, mode, fallback=False)


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, FIN_TILE - 1, FIN_TILE, FIN_TILE + 1, FIN_TILE + FIN_LIMIT,
                               FIN_TILE + FIN_LIMIT + 1])
def test_small_and_ragged_sizes(n):
    rng = np.random.default_rng(n)
    for mode in (2, 3):
        keys = _segments(n, 5, 8 * mode, rng)
        _check(keys, mode, fallback=False)
        keys = rng.integers(0, 1 << 63, n, dtype=np.uint64)
        _check(keys, mode)


@pytest.mark.parametrize("kind", ["uniform", "terrain", "urban", "skewed"])
def test_tiling_result_is_the_same_in_every_sort_mode(port_oracle, kind):
    """The whole path (index, sort, sweep) on a cloud, every sort mode against the oracle."""
    _torch_cuda()
    import schwarzwald_b200 as sw
    from oracle import sworacle
    from schwarzwald_b200 import synth
    n = 400_000
    xyz = synth.generate(kind, n, 7, device="cpu").numpy()
    xyz[1000:1400] = xyz[1000]  # duplicates: a run of 400 identical keys
    bmin, bmax = sw.cubic_bounds(xyz.min(0), xyz.max(0))
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax, 250.0)
    params = sworacle.make_params("RANDOM_GRID", "FAST", spacing, bmin, bmax, max_points_per_node=2000, concurrency=8)
    want = port_oracle.tile(params, xyz)
    wt, wi = want.canonical()
    for mode in (-1, 0, 1, 2, 3):
        with sw.GpuTiler("RANDOM_GRID", "FAST", bmin, bmax, spacing, max_points_per_node=2000, concurrency=8) as t:
            t.set_sort_mode(mode)
            for _ in range(2):  # the second batch of the automatic mode follows the first one's run lengths
                got = t.tile(xyz.copy())
                keys, order = t.keys(n)
                assert np.array_equal(keys, want.keys) and np.array_equal(order, want.order), (kind, mode)
                gt, gi = got.canonical()
                assert np.array_equal(wt[:, :3], gt[:, :3]) and np.array_equal(wi, gi), (kind, mode)
            st = t.stats()
            if mode > 0:  # clustered clouds hold runs the finish kernel leaves to the eight-pass fallback
                assert st["sort_passes"] == 8 - mode + (8 if st["sort_fallback"] == 1 else 0)
                assert st["sort_fallback"] == 0 or kind == "urban"
