#!/usr/bin/env python
"""Generates tests/golden/batches_golden.json from oracle/_ref: multi-batch tiling with the
reference's own primitives (index_point, calculate_morton_index relative to node bounds, sample_points, ...)
compiled verbatim from /root/reference and the orchestration of oracle/orchestrator.h.  Digests only; committed so
that a box without the reference tree can still pin against it."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SAMPLINGS = ["RANDOM_GRID", "GRID_CENTER", "JITTERED", "MIN_DISTANCE", "MIN_DISTANCE_FAST"]
SPLITS = {"ACCURATE": [20_000, 45_000, 25_000], "FAST": [40_000, 30_000, 20_000]}


def case_input():
    import schwarzwald_b200 as sw
    rng = np.random.default_rng(33)
    xyz = np.round(rng.random((90_000, 3)) * np.array([400.0, 300.0, 60.0]) + np.array([1000.0, -20.0, 5.0]), 3)
    xyz[:5] -= 700.0  # outliers of the first batch get clamped
    bmin, bmax = sw.cubic_bounds(xyz[5:].min(0), xyz[5:].max(0))
    return xyz, bmin, bmax, sw.spacing_from_diagonal_fraction(bmin, bmax)


def digest(res):
    table, ids = res.canonical()
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(table).tobytes())
    h.update(np.ascontiguousarray(ids).tobytes())
    return h.hexdigest()


def run(oracle, sampling, tiling):
    from oracle import sworacle
    xyz, bmin, bmax, spacing = case_input()
    p = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=800, concurrency=2)
    return oracle.tile_batches(p, xyz, SPLITS[tiling])


def main():
    from oracle import sworacle
    ref = sworacle.Oracle("ref")
    cases = []
    for tiling in ("ACCURATE", "FAST"):
        for sampling in SAMPLINGS:
            res = run(ref, sampling, tiling)
            cases.append({"sampling": sampling, "tiling": tiling, "batches": SPLITS[tiling], "nodes": int(len(res.nodes)),
                          "ids": int(len(res.ids)), "start_level": res.start_level, "digest": digest(res)})
    out = {"generator": "tests/golden/make_golden_batches.py",
           "source": "oracle/_ref/libswref.so (reference primitives verbatim, orchestration of oracle/orchestrator.h)",
           "cases": cases}
    with open(os.path.join(HERE, "batches_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
