"""The five BASELINE.json configurations as data: generator, size, sampling, tiling.

Shared by bench.py and the full-size parity tests.  Input generation is not part of the tiler hot path.
"""
from __future__ import annotations

import numpy as np

from . import synth
from .tiler import cubic_bounds, cubic_bounds_at_origin, spacing_from_diagonal_fraction

MAX_POINTS_PER_NODE = 20000

CONFIGS = {
    # configs[0]: synthetic uniform 10M-point LAS, --tiler GRID_CENTER FAST -> 3DTILES (runs on the CPU reference)
    "c1": dict(points=10_000_000, kind="uniform", seed=1, gen_kw=dict(side_m=1000.0), sampling="GRID_CENTER",
               tiling="FAST", concurrency=8, shift_float32=True,
               workload="synthetic uniform 10M-point cloud, GRID_CENTER FAST, 3DTILES pre-transform "
                        "(shift to centre + float32), single batch"),
    # configs[1]: synthetic 100M-point terrain-like cloud, RANDOM_GRID FAST, single B200
    "c2": dict(points=100_000_000, kind="terrain", seed=2, gen_kw={}, sampling="RANDOM_GRID", tiling="FAST",
               concurrency=32, shift_float32=False,
               workload="synthetic 100M-point terrain-like cloud, RANDOM_GRID FAST, single batch"),
    # configs[2]: synthetic 500M-point clustered urban-LiDAR-like cloud, JITTERED FAST, 1/2/4/8 B200
    "c3": dict(points=500_000_000, kind="urban", seed=3, gen_kw={}, sampling="JITTERED", tiling="FAST",
               concurrency=32, shift_float32=False,
               workload="synthetic 500M-point clustered urban-LiDAR-like cloud, JITTERED FAST, single batch"),
    # configs[3]: synthetic 1B-point cloud, MIN_DISTANCE ACCURATE, Morton-prefix sharded across 8 B200
    "c4": dict(points=1_000_000_000, kind="terrain", seed=4, gen_kw=dict(side_m=30000.0), sampling="MIN_DISTANCE",
               tiling="ACCURATE", concurrency=32, shift_float32=False,
               workload="synthetic 1B-point terrain-like cloud (30 km), MIN_DISTANCE ACCURATE, single batch"),
    # configs[4]: skewed density stress (95% of points in 1% of volume), 2B points, GRID_CENTER ENTWINE_LAZ
    # (the CLI's default tiling strategy is FAST, executable/main.cpp:300-301)
    "c5": dict(points=2_000_000_000, kind="skewed", seed=5, gen_kw={}, sampling="GRID_CENTER", tiling="FAST",
               concurrency=32, shift_float32=False,
               workload="synthetic 2B-point skewed cloud (95% of points in 1% of the volume), GRID_CENTER FAST, "
                        "single batch"),
}


def default_config(n_gpus):
    """bench.py without --config: BASELINE's single-GPU config on one GPU, its 1/2/4/8-GPU config otherwise."""
    return "c2" if n_gpus <= 1 else "c3"


def generate_slice(cfg, start, count, device, chunk=1 << 24):
    """Points [start, start + count) of the config's cloud as a (count, 3) float64 torch tensor on `device`."""
    import torch
    gen = synth.GENERATORS[cfg["kind"]]
    out = torch.empty((count, 3), dtype=torch.float64, device=device)
    for s in range(0, count, chunk):
        m = min(chunk, count - s)
        out[s:s + m] = gen(m, seed=cfg["seed"], device=device, start=start + s, **cfg["gen_kw"])
    return out


def finish_bounds(cfg, tight_min, tight_max):
    """Cubic bounds + spacing of a config from the tight bounds of the full cloud, as the tiler CLI derives them
    (AABB::makeCubic, math/AABB.h:50-61; 3DTILES: cubic bounds re-centred at the origin, Tiler.cpp:185-187;
    spacing = diagonal / 250, TilerProcess.cpp:598-604).  Returns (bmin, bmax, spacing, centre or None)."""
    tight_min = np.asarray(tight_min, np.float64)
    tight_max = np.asarray(tight_max, np.float64)
    if cfg["shift_float32"]:
        cmin, cmax = cubic_bounds(tight_min, tight_max)
        centre = cmin + (cmax - cmin) / 2
        bmin, bmax = cubic_bounds_at_origin(tight_min, tight_max)
    else:
        bmin, bmax = cubic_bounds(tight_min, tight_max)
        centre = None
    return bmin, bmax, spacing_from_diagonal_fraction(bmin, bmax), centre


def apply_pre_transform(cfg, xyz, centre):
    """The 3DTILES pre-transform of C1 (process/TilerProcess.cpp:552-559): p -= centre; p = (double)(float)p."""
    if not cfg["shift_float32"]:
        return xyz
    import torch
    c = torch.tensor(np.asarray(centre, np.float64), dtype=torch.float64, device=xyz.device)
    return (xyz - c).to(torch.float32).to(torch.float64)
