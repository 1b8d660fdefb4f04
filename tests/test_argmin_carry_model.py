"""A Python model of how select_argmin_kernel + argmin_carry_kernel (schwarzwald_b200/csrc/kernels_sampling.cu) settle
selection cells that cross tile boundaries: every tile works alone, leaves the partial minimum of its first cell in a
descriptor (`first`) when that cell ends inside the tile, and the minimum of the run that is open at its end (`agg`,
with a flag whether the tile holds a cell head at all); one thread per tile then walks back over the `agg` entries.
Checked against a plain first-minimum per cell on random head patterns with tiny tiles (cells that span many tiles,
tiles without any head, ties).  The CUDA code itself is checked on the GPU (tests/test_gpu_parity.py).
"""
import random

import pytest


def argmin_op(earlier, later):  # strict: ties keep the earlier point
    return later if later[0] < earlier[0] else earlier


def model(cell, d, tile):
    n = len(cell)
    head = [i == 0 or cell[i] != cell[i - 1] for i in range(n)]
    tail = [i == n - 1 or cell[i] != cell[i + 1] for i in range(n)]
    tiles = (n + tile - 1) // tile
    sel = [0] * n
    first = [None] * tiles
    agg = [None] * tiles
    for t in range(tiles):  # select_argmin_kernel: tiles are independent
        lo, hi = t * tile, min(n, (t + 1) * tile)
        from_tile_start, cur, seen_head = True, None, False
        for i in range(lo, hi):
            v = (d[i], i)
            if head[i]:
                from_tile_start = False
                seen_head = True
            cur = v if (head[i] or cur is None) else argmin_op(cur, v)
            if tail[i]:
                if from_tile_start and t != 0:
                    first[t] = cur
                else:
                    sel[cur[1]] = 1
        agg[t] = (cur, seen_head)  # the run that is open at the end of the tile
    for t in range(tiles):  # argmin_carry_kernel: one thread per tile
        if first[t] is None:
            continue
        best = first[t]
        u = t - 1
        while u >= 0:
            e, has_head = agg[u]
            best = argmin_op(e, best)
            if has_head:
                break
            u -= 1
        sel[best[1]] = 1
    return sel


@pytest.mark.parametrize("tile", [1, 2, 3, 8, 16])
def test_carry_model_selects_the_first_minimum_of_every_cell(tile):
    rng = random.Random(tile)
    for _ in range(400):
        n = rng.randint(1, 12 * tile + 3)
        cell, c = [], 0
        while len(cell) < n:
            length = rng.choice([1, 1, 2, 3, tile, tile + 1, 5 * tile, rng.randint(1, 3 * tile + 1)])
            cell += [c] * min(length, n - len(cell))
            c += 1
        d = [rng.choice([0.0, 1.0, 2.0, rng.random()]) for _ in range(n)]
        want = [0] * n
        i = 0
        while i < n:
            e = i
            while e < n and cell[e] == cell[i]:
                e += 1
            best = min(range(i, e), key=lambda q: (d[q], q))
            want[best] = 1
            i = e
        assert model(cell, d, tile) == want
