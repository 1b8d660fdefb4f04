"""A Python model of segment_finish_kernel (schwarzwald_b200/csrc/kernels_index_sort.cu), statement by statement where
it matters: tile ownership of runs, the warp windows of the 40-bit mode, the scalar scan, the long-run list.  The model
is fuzzed against sorted() with SMALL tile / limit parameters, so that the corner cases (runs that end exactly at a
tile or window boundary, runs of exactly LIMIT elements, foreign runs, ragged last tiles) occur thousands of times —
far more often than in the GPU tests, which use the kernel's real 4 096 / 256.  Tiles are processed in random order on
the live arrays: a tile may see keys and ids that other tiles have already permuted, as on the GPU.

This checks the ALGORITHM (the part of the kernel that is not covered by a sanitizer); the CUDA code itself is checked
on the GPU by tests/test_gpu_sort_modes.py.
"""
import random

import pytest

HEAD = 1 << 31
FOREIGN = 1 << 30
KEEP = None
STRIDE, LANES = 24, 32


def finish_model(keys, ids, low_bits, tile, limit, windows, rng):
    """Returns (unsorted_long, long_elements, long_runs); keys / ids are permuted in place."""
    n = len(keys)
    mask = (1 << low_bits) - 1
    window = tile + limit
    unsorted_long, long_elements, long_runs = False, 0, []
    order = list(range((n + tile - 1) // tile))
    rng.shuffle(order)
    for t in order:
        base = t * tile
        valid = min(n - base, window)
        # ---- tags ----
        tag = [0] * (1 + window + LANES)
        for j in range(window):
            g = base + j
            key = keys[g] if j < valid else (1 << 64) - 1
            pk = keys[g - 1] if (g > 0 and j <= valid) else (1 << 64) - 1
            head = ((key ^ pk) >> low_bits) != 0 or g == 0 or j >= valid
            tag[1 + j] = (key & mask) | (HEAD if head else 0)
            if j == 0:
                tag[0] = (pk & mask) | HEAD | FOREIGN
        for j in range(window, window + LANES):
            tag[1 + j] = HEAD
        src = [KEEP] * window

        def scan_run(j):
            nonlocal unsorted_long, long_elements
            tg = tag[1 + j]
            lo = tg & mask
            budget = limit - 1
            too_long = False
            l, tt, rk = j, tg, 0
            while not (tt & HEAD):
                if budget == 0:
                    too_long = True
                    break
                budget -= 1
                l -= 1
                tt = tag[1 + l]
                rk += 1 if (tt & mask) <= lo else 0
            skip = False
            if not too_long:
                skip = bool(tt & FOREIGN) or l >= tile
                r = j + 1
                while not skip:
                    t2 = tag[1 + r]
                    if t2 & HEAD:
                        break
                    if budget == 0:
                        too_long = True
                        break
                    budget -= 1
                    rk += 1 if (t2 & mask) < lo else 0
                    r += 1
            if too_long:
                if not (tg & HEAD) and (tag[j] & mask) > lo:
                    unsorted_long = True
                if j < tile:
                    long_elements += 1
                if tg & HEAD:
                    long_runs.append(base + j)
            elif not skip:
                p = l + rk
                if p != j:
                    assert src[p] is KEEP
                    src[p] = j

        if not windows:
            for j in range(valid):
                if not (tag[1 + j] & tag[2 + j] & HEAD):
                    scan_run(j)
        else:
            w0 = 0
            while w0 < valid:
                tags = [tag[1 + w0 + lane] for lane in range(LANES)]
                heads = [bool(x & HEAD) for x in tags]
                for lane in range(LANES):
                    j = w0 + lane
                    live = j < valid
                    hb = [i for i in range(lane + 1) if heads[i]]
                    ha = [i for i in range(lane + 1, LANES) if heads[i]]
                    l_lane = hb[-1] if hb else -1
                    r_lane = ha[0] if ha else LANES
                    mine = live and bool(hb) and l_lane < STRIDE
                    fast = mine and bool(ha)
                    slow = live and ((mine and not ha) or (not hb and (lane >= LANES - STRIDE or w0 == 0)))
                    if fast:
                        val = ((tags[lane] & mask) << 5) | lane
                        rank = sum(1 for q in range(l_lane, r_lane)
                                   if q != lane and (((tags[q] & mask) << 5) | q) < val)
                        l = w0 + l_lane
                        p = l + rank
                        if l < tile and p != j:
                            assert src[p] is KEEP
                            src[p] = j
                    if slow:
                        scan_run(j)
                w0 += STRIDE
        # ---- pull ----
        new_ids = [ids[base + s] if s is not KEEP else None for s in src]
        for j in range(window):
            s = src[j]
            if s is not KEEP:
                g = base + j
                keys[g] = (keys[g] & ~mask) | (tag[1 + s] & mask)
                ids[g] = new_ids[j]
    return unsorted_long, long_elements, long_runs


def make_keys(rng, n, low_bits, run_len):
    """Sorted by the top bits (ids ascending inside a run, as the stable passes leave them), random low bits."""
    keys, hi = [], 1
    while len(keys) < n:
        length = run_len(rng)
        lo_range = rng.choice([2, 5, 1 << low_bits])
        for _ in range(min(length, n - len(keys))):
            keys.append((hi << low_bits) | rng.randrange(min(lo_range, 1 << low_bits)))
        hi += rng.choice([1, 1, 2, 1 << 20])
    return keys


def check(rng, n, low_bits, tile, limit, windows, run_len):
    keys = make_keys(rng, n, low_bits, run_len)
    ids = sorted(rng.sample(range(10 * n + 10), n))  # ascending inside every run
    orig = list(zip(keys, ids))
    unsorted_long, long_elements, long_runs = finish_model(keys, ids, low_bits, tile, limit, windows, rng)
    # expected: every run of at most `limit` elements ordered by (low bits, id); longer runs untouched and listed
    i, exp, exp_long, exp_unsorted, exp_long_elements = 0, [], [], False, 0
    while i < n:
        e = i
        while e < n and (orig[e][0] >> low_bits) == (orig[i][0] >> low_bits):
            e += 1
        run = orig[i:e]
        if e - i > limit:
            exp += run
            exp_long.append(i)
            exp_long_elements += e - i
            exp_unsorted |= any(run[q][0] > run[q + 1][0] for q in range(len(run) - 1))
        else:
            exp += sorted(run)
        i = e
    assert list(zip(keys, ids)) == exp
    assert sorted(long_runs) == exp_long
    assert unsorted_long == exp_unsorted
    # the count is a lower bound used by a heuristic only (elements the previous tile reaches are not counted twice)
    assert long_elements <= exp_long_elements and (exp_long_elements == 0) == (long_elements == 0)


# (the warp windows rank runs of up to 32 elements without looking at the limit, so the limit must not be smaller
# than a window: the kernel has 256)
@pytest.mark.parametrize("windows,tile,limit", [(False, 64, 16), (False, 48, 40), (False, 96, 8),
                                                (True, 64, 40), (True, 48, 32), (True, 96, 33), (True, 72, 64)])
def test_model_matches_sorted(windows, tile, limit):
    rng = random.Random(1000 * tile + limit + int(windows))
    shapes = [
        lambda r: 1,
        lambda r: r.choice([1, 1, 1, 2, 3]),
        lambda r: r.randint(1, 12),
        lambda r: r.choice([1, 2, limit - 1, limit, limit + 1]),
        lambda r: r.choice([1, 3, 3 * tile, limit + 2]),
        lambda r: r.randint(1, 2 * limit),
    ]
    for round_ in range(150):
        n = rng.choice([1, 2, tile - 1, tile, tile + 1, tile + limit, tile + limit + 1, 3 * tile + 5,
                        rng.randint(1, 6 * tile)])
        check(rng, n, rng.choice([3, 8]), tile, limit, windows, shapes[round_ % len(shapes)])
