"""Multi-batch mode at the reference's default batch size: N points of a BASELINE cloud in batches of 10 M
(executable/main.cpp:233-236) through swgpu_set_multi_batch; wall time per batch and in total, device-resident input.
usage: python tools/bench_batches.py [config] [points] [batch] [--check]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import schwarzwald_b200 as sw
from schwarzwald_b200 import workloads

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000
check = "--check" in sys.argv
cfg = workloads.CONFIGS[cfg_name]
dev = torch.device("cuda", 0)
mn, mx, xyz = bench.full_cloud_tight_bounds(cfg, n, dev, keep=(0, n))
bmin, bmax, spacing, centre = workloads.finish_bounds(cfg, mn, mx)
xyz = workloads.apply_pre_transform(cfg, xyz, centre)
t = sw.GpuTiler(cfg["sampling"], cfg["tiling"], bmin, bmax, spacing, max_points_per_node=workloads.MAX_POINTS_PER_NODE,
                concurrency=cfg["concurrency"])
t.set_multi_batch(True)
torch.cuda.synchronize()
t0 = time.perf_counter()
per_batch = []
for lo in range(0, n, batch):
    hi = min(n, lo + batch)
    tb = time.perf_counter()
    t.build_execution_graph(xyz[lo:hi].contiguous() if lo % 2 else xyz[lo:hi])
    torch.cuda.synchronize()
    per_batch.append((time.perf_counter() - tb) * 1e3)
tf = time.perf_counter()
t.finalize()
torch.cuda.synchronize()
t1 = time.perf_counter()
nn, ni = t.result_size()
print("config %s %d points in batches of %d: total %.1f ms (%.2f G points/s), finalize %.1f ms, nodes %d ids %d"
      % (cfg_name, n, batch, (t1 - t0) * 1e3, n / (t1 - t0) / 1e9, (t1 - tf) * 1e3, nn, ni))
print("per batch ms:", " ".join("%.1f" % x for x in per_batch))
if check:
    from oracle import sworacle
    res = t.result()
    orc = sworacle.Oracle("ref" if sworacle.have_ref() else "port")
    p = sworacle.make_params(cfg["sampling"], cfg["tiling"], spacing, bmin, bmax,
                             max_points_per_node=workloads.MAX_POINTS_PER_NODE, concurrency=cfg["concurrency"])
    sizes = [min(batch, n - lo) for lo in range(0, n, batch)]
    tw = time.perf_counter()
    want = orc.tile_batches(p, xyz.cpu().numpy(), sizes)
    print("oracle %.1f s" % (time.perf_counter() - tw))
    wt, wi = want.canonical()
    gt, gi = res.canonical()
    print("PARITY", "OK" if np.array_equal(wt, gt) and np.array_equal(wi, gi) else "MISMATCH", len(wt), len(gt))
t.close()
