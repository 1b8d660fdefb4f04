#!/usr/bin/env python
"""Generates tests/golden/tiler_golden.json from oracle/_ref (the reference's own hot-path code
compiled verbatim from /root/reference).  Run here, where /root/reference exists; the JSON is
committed so that the GPU box (which has no reference tree) can still pin against the reference.

Each case stores a digest of the full result (node table + per-node point ids + sorted keys).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CLOUDS = {
    "uniform_40k": ("uniform", 40_000, 101, {"side_m": 80.0}),
    "terrain_90k": ("terrain", 90_000, 102, {"side_m": 600.0}),
}
SAMPLINGS = ["RANDOM_GRID", "GRID_CENTER", "JITTERED", "MIN_DISTANCE"]
TILINGS = ["ACCURATE", "FAST"]
# strategies added after the first fixture set: appended so that earlier case numbers stay put
LATER_SAMPLINGS = ["MIN_DISTANCE_FAST"]


def case_input(cloud):
    import schwarzwald_b200 as sw
    from schwarzwald_b200 import synth
    kind, n, seed, kw = CLOUDS[cloud]
    xyz = synth.generate(kind, n, seed, device="cpu", **kw).numpy()
    bmin, bmax = sw.cubic_bounds(xyz.min(0), xyz.max(0))
    return xyz, bmin, bmax, sw.spacing_from_diagonal_fraction(bmin, bmax)


def digest(res):
    table, ids = res.canonical()
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(table).tobytes())
    h.update(np.ascontiguousarray(ids).tobytes())
    h.update(np.ascontiguousarray(res.keys).tobytes())
    h.update(np.ascontiguousarray(res.order).tobytes())
    return h.hexdigest()


def main():
    from oracle import sworacle
    ref = sworacle.Oracle("ref")
    cases = []
    for group in (SAMPLINGS, LATER_SAMPLINGS):
      for cloud in CLOUDS:
        xyz, bmin, bmax, spacing = case_input(cloud)
        for tiling in TILINGS:
            for sampling in group:
                p = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=700, concurrency=2)
                res = ref.tile(p, xyz)
                cases.append({"cloud": cloud, "sampling": sampling, "tiling": tiling, "max_points": 700,
                              "concurrency": 2, "nodes": int(len(res.nodes)), "ids": int(len(res.ids)),
                              "start_level": res.start_level, "digest": digest(res)})
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libswref.so (reference TUs, verbatim)",
           "cases": cases}
    with open(os.path.join(HERE, "tiler_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
