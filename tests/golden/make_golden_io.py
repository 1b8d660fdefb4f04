#!/usr/bin/env python
"""Generates tests/golden/io_golden.json from oracle/_ref: the reference's own LAS reader conversion
(io/LASFile.cpp), PNTS position attribute (io/PNTSWriter.cpp) and LAS persistence (io/LASPersistence.*)
compiled verbatim from /root/reference, run on seeded inputs.  The JSON stores SHA-256 digests (plus a
few literal values) and is committed, so the GPU box can pin against the reference without its tree.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# LAS headers of the synthetic files: (scale, offset, header min, header max)
LAS_CASES = {
    "utm_mm": ([0.001, 0.001, 0.001], [500000.0, 4200000.0, 0.0], [500000.25, 4200010.5, -3.0],
               [502100.75, 4202050.125, 410.5]),
    # node diagonals cross the 1 m threshold of compute_las_scale_from_bounds a few levels down
    "small_object": ([0.0001, 0.0001, 0.0001], [10.0, -2.0, 0.5], [10.0, -2.0, 0.5], [11.5, -0.75, 1.875]),
    # root diagonal above 1 000 km (scale 0.01), children below
    "continental": ([0.01, 0.01, 0.01], [0.0, 0.0, 0.0], [1000.0, 2000.0, -50.0], [701000.0, 650000.0, 4200.0]),
    "aniso_negative": ([0.01, 0.0025, 0.0001], [-1250.5, 33.125, -80.0], [-1300.0, 20.0, -80.5], [900.0, 5000.0, 12.0]),
}


def las_input(name, n=60_000, seed=7):
    """Seeded LAS record coordinates; ~2 % of the records fall outside the header bounds (the reader clamps
    them), a few sit at the int32 extremes."""
    scale, offset, hmin, hmax = (np.array(v, np.float64) for v in LAS_CASES[name])
    rng = np.random.default_rng(seed)
    lo = np.floor((hmin - offset) / scale).astype(np.int64)
    hi = np.ceil((hmax - offset) / scale).astype(np.int64)
    span = hi - lo
    las = lo + (rng.random((n, 3)) * (span * 1.02) - span * 0.01).astype(np.int64)
    las[:4] = [[2**31 - 1] * 3, [-2**31] * 3, [0, 0, 0], [-1, 1, -1]]
    return np.clip(las, -2**31, 2**31 - 1).astype(np.int32), scale, offset, hmin, hmax


def cubic(hmin, hmax):
    import schwarzwald_b200 as sw
    return sw.cubic_bounds(hmin, hmax)


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def tiling_case(oracle, name, shift, sampling="GRID_CENTER", tiling="ACCURATE", max_points=500):
    """LAS records -> positions -> tiling -> payloads, everything from `oracle`."""
    import schwarzwald_b200 as sw
    from oracle import sworacle
    las, scale, offset, hmin, hmax = las_input(name)
    cmin, cmax = cubic(hmin, hmax)
    center = cmin + (cmax - cmin) / 2 if shift else None
    t = sworacle.make_las_transform(scale, offset, hmin, hmax, center)
    xyz = oracle.las_positions(las, t)
    bmin, bmax = (cmin - center, cmax - center) if shift else (cmin, cmax)
    # coarse spacing for the small object: a deeper tree, so that node diagonals drop below 1 m
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax, 12.0 if name == "small_object" else 250.0)
    p = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=max_points, concurrency=2)
    res, clamped = oracle.tile(p, xyz, return_clamped=True)
    pnts = oracle.payload_pnts(clamped, res.ids)
    las_out, headers = oracle.payload_las(clamped, res.ids, res.nodes, (bmin, bmax))
    return dict(las=las, transform=t, xyz=xyz, clamped=clamped, bounds=(bmin, bmax), spacing=spacing, params=p,
                result=res, pnts=pnts, las_out=las_out, headers=headers)


def canonical_payload(res, payload):
    """payload rows re-ordered like TileResult.canonical() orders the ids"""
    order = np.lexsort((res.nodes["index"], res.nodes["levels"]))
    chunks = [payload[int(n["first"]): int(n["first"]) + int(n["count"])] for n in res.nodes[order]]
    return np.concatenate(chunks) if chunks else payload[:0]


def canonical_headers(res, headers):
    return headers[np.lexsort((res.nodes["index"], res.nodes["levels"]))]


def main():
    from oracle import sworacle
    ref = sworacle.Oracle("ref")
    cases = []
    for name in LAS_CASES:
        for shift in (False, True):
            c = tiling_case(ref, name, shift)
            cases.append({
                "las": name, "shift": shift,
                "positions": sha(c["xyz"]), "first_positions": c["xyz"][:6].tolist(),
                "nodes": int(len(c["result"].nodes)), "ids": int(len(c["result"].ids)),
                "pnts": sha(canonical_payload(c["result"], c["pnts"])),
                "las_records": sha(canonical_payload(c["result"], c["las_out"])),
                "las_headers": sha(canonical_headers(c["result"], c["headers"])),
                "scales": sorted(set(float(s) for s in c["headers"]["scale"])),
            })
    out = {"generator": "tests/golden/make_golden_io.py",
           "source": "oracle/_ref/libswref.so (reference io/LASFile.cpp, io/PNTSWriter.cpp, io/LASPersistence.* verbatim; "
                     "LASzip's coordinate quantisation restated in oracle/shim/laszip_api.h)",
           "cases": cases}
    with open(os.path.join(HERE, "io_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
