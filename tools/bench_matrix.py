#!/usr/bin/env python
"""Stage times of the hot path for every sampling strategy x tiling mode on the BASELINE cloud shapes
(device-resident input, CUDA events from the library).  Not the bench line: exploration / DESIGN.md table.

    python tools/bench_matrix.py [--points 100000000] [--out gpurun_out/matrix.json] [--cases a,b,...]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, cloud, sampling, tiling): the BASELINE.json configs at single-GPU size + the remaining strategies
CASES = [
    ("c0_uniform_grid_center_fast", "uniform", "GRID_CENTER", "FAST"),
    ("c1_terrain_random_grid_fast", "terrain", "RANDOM_GRID", "FAST"),
    ("c2_urban_jittered_fast", "urban", "JITTERED", "FAST"),
    ("c3_terrain_min_distance_accurate", "terrain", "MIN_DISTANCE", "ACCURATE"),
    ("c4_skewed_grid_center_fast", "skewed", "GRID_CENTER", "FAST"),
    ("x_terrain_random_grid_accurate", "terrain", "RANDOM_GRID", "ACCURATE"),
    ("x_terrain_min_distance_fast_fast", "terrain", "MIN_DISTANCE_FAST", "FAST"),
    ("x_urban_min_distance_fast", "urban", "MIN_DISTANCE", "FAST"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=100_000_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "matrix.json"))
    ap.add_argument("--cases", default="")
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    import torch

    import schwarzwald_b200 as sw
    from schwarzwald_b200 import synth

    dev = torch.device("cuda", 0)
    want = set(args.cases.split(",")) if args.cases else None
    rows = []
    clouds = {}
    for name, cloud, sampling, tiling in CASES:
        if want and name not in want:
            continue
        if cloud not in clouds:
            clouds.clear()  # one cloud resident at a time
            gen = {"uniform": synth.uniform_cube, "terrain": synth.terrain, "urban": synth.urban,
                   "skewed": synth.skewed}[cloud]
            xyz = torch.empty((args.points, 3), dtype=torch.float64, device=dev)
            chunk = 1 << 24
            for s in range(0, args.points, chunk):
                m = min(chunk, args.points - s)
                xyz[s:s + m] = gen(m, device=dev, start=s)
            clouds[cloud] = xyz
        xyz = clouds[cloud]
        mn, mx = synth.tight_bounds(xyz)
        bmin, bmax = sw.cubic_bounds(mn, mx)
        spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
        row = {"case": name, "cloud": cloud, "sampling": sampling, "tiling": tiling, "points": args.points}
        try:
            with sw.GpuTiler(sampling, tiling, bmin, bmax, spacing, concurrency=32) as t:
                t.set_stream(torch.cuda.current_stream().cuda_stream)
                t.enable_timing(True)
                best = None
                for it in range(1 + args.steps):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    t.build_execution_graph(xyz)
                    t.finalize()
                    torch.cuda.synchronize()
                    wall = (time.perf_counter() - t0) * 1e3
                    st = t.stats()
                    if it > 0 and (best is None or wall < best["wall_ms"]):
                        best = {"wall_ms": wall, **{k: st[k] for k in st}}
                row.update({k: (round(v, 3) if isinstance(v, float) else v) for k, v in best.items()})
                row["points_per_s"] = args.points / (best["wall_ms"] * 1e-3)
                row["start_level"] = t.start_level()
                row["peak_mem_gb"] = round(torch.cuda.max_memory_allocated() / 1e9, 2)
                free, total = torch.cuda.mem_get_info()
                row["device_mem_used_gb"] = round((total - free) / 1e9, 2)
        except Exception as e:
            row["error"] = repr(e)
        print(json.dumps(row), flush=True)
        rows.append(row)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
