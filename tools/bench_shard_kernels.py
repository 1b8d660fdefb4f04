#!/usr/bin/env python
"""Times the shard-side kernels (prefix histogram, partition) on one GPU with CUDA events."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import schwarzwald_b200 as sw  # noqa: E402
from schwarzwald_b200 import distributed, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda", 0)
xyz = torch.empty((n, 3), dtype=torch.float64, device=dev)
for s in range(0, n, 1 << 24):
    m = min(1 << 24, n - s)
    xyz[s:s + m] = synth.terrain(m, seed=2, device=dev, start=s)
mn, mx = synth.tight_bounds(xyz)
bmin, bmax = sw.cubic_bounds(mn, mx)
spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
t = sw.GpuTiler("RANDOM_GRID", "FAST", bmin, bmax, spacing, concurrency=32)
lib = t._lib
keys = torch.empty(n, dtype=torch.int64, device=dev)
bins = torch.zeros(262144, dtype=torch.int32, device=dev)


def timed(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-40s %8.3f ms" % (name, e0.elapsed_time(e1) / reps), flush=True)


timed("morton_encode", lambda: t.morton_encode_device(xyz.data_ptr(), n, keys.data_ptr()))
timed("prefix_histogram", lambda: t._check(lib.swgpu_prefix_histogram_device(
    t._h, C.c_void_p(keys.data_ptr()), n, C.c_void_p(bins.data_ptr()))))
hb = (bins.cpu().numpy().view(np.uint32) // 4).astype(np.uint32)
out_xyz = torch.empty_like(xyz)
out_id = torch.empty(n, dtype=torch.int32, device=dev)
for world in (2, 8):
    fp = distributed.choose_splitters(hb, world, 6)
    sc = np.zeros(world, np.uint64)
    timed("partition world=%d" % world, lambda: t._check(lib.swgpu_partition_device(
        t._h, C.c_void_p(keys.data_ptr()), C.c_void_p(xyz.data_ptr()), n, C.c_void_p(fp.ctypes.data), world, 0,
        C.c_void_p(out_xyz.data_ptr()), C.c_void_p(out_id.data_ptr()), C.c_void_p(sc.ctypes.data))))
    print("   send counts", sc.tolist())
timed("torch copy 2.4 GB", lambda: out_xyz.copy_(xyz))
