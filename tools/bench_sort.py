#!/usr/bin/env python
"""Times the radix sort alone (swgpu_sort_keys_device: key histogram + passes) on the Morton keys of the bench
cloud.  Used with SWGPU_LIB=<variant .so> to A/B kernel variants and the RS_EXPERIMENT phase-removal builds."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import schwarzwald_b200 as sw  # noqa: E402
from schwarzwald_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda", 0)
xyz = torch.empty((n, 3), dtype=torch.float64, device=dev)
for s in range(0, n, 1 << 24):
    m = min(1 << 24, n - s)
    xyz[s:s + m] = synth.terrain(m, seed=2, device=dev, start=s)
mn, mx = synth.tight_bounds(xyz)
bmin, bmax = sw.cubic_bounds(mn, mx)
t = sw.GpuTiler("RANDOM_GRID", "FAST", bmin, bmax, sw.spacing_from_diagonal_fraction(bmin, bmax))
keys0 = torch.empty(n, dtype=torch.int64, device=dev)
t.morton_encode_device(xyz.data_ptr(), n, keys0.data_ptr())
del xyz
keys = torch.empty_like(keys0)
order = torch.empty(n, dtype=torch.int32, device=dev)
times = []
for it in range(4):
    keys.copy_(keys0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t.sort_keys_device(keys.data_ptr(), n, order.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    if it:
        times.append(e0.elapsed_time(e1))
ok = bool((keys[1:] >= keys[:-1]).all().item())
print("%-28s sort_keys_device %.3f ms (min %.3f)  sorted=%s" % (os.path.basename(os.environ.get("SWGPU_LIB", "libswgpu.so")),
                                                           sum(times) / len(times), min(times), ok), flush=True)
