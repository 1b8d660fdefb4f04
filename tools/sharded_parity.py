#!/usr/bin/env python
"""Multi-rank parity check (run under torchrun, one rank per GPU):
every rank tiles its slice through ShardedTiler (NCCL all-reduce + all-to-all over NVLink), the
per-rank results are gathered on rank 0, merged, and compared bit for bit with a single-GPU run
of the same cloud.  Prints one 'PARITY OK ...' line per case on rank 0; exits non-zero on mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/sharded_parity.py [--points 4000000]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import schwarzwald_b200 as sw  # noqa: E402
from schwarzwald_b200 import distributed, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=4_000_000, help="points in total")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = args.points
    cuts = np.linspace(0, n, world + 1).astype(int)
    ok = True
    for kind, sampling, tiling in (("terrain", "RANDOM_GRID", "FAST"), ("terrain", "GRID_CENTER", "ACCURATE"),
                                   ("urban", "JITTERED", "FAST"), ("skewed", "GRID_CENTER", "FAST")):
        full = synth.generate(kind, n, 3, device=dev)
        mn, mx = synth.tight_bounds(full)
        bmin, bmax = sw.cubic_bounds(mn, mx)
        spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
        mine = full[cuts[rank]:cuts[rank + 1]].clone()
        with distributed.ShardedTiler(sampling, tiling, bmin, bmax, spacing, concurrency=8, device=local) as st:
            st.set_stream(torch.cuda.current_stream().cuda_stream)
            st.build_execution_graph(mine)
            st.finalize()
            res = st.result()
            info = dict(st.last)
        parts = [None] * world if rank == 0 else None
        dist.gather_object((res.nodes, res.ids, res.start_level, info["n_shard"]), parts, dst=0)
        if rank == 0:
            merged = distributed.merge_results([sw.TileResult(p[0], p[1], p[2]) for p in parts])
            with sw.GpuTiler(sampling, tiling, bmin, bmax, spacing, concurrency=8, device=local) as t:
                want = t.tile(full)
            wt, wi = want.canonical()
            gt, gi = merged.canonical()
            same = (np.array_equal(wt[:, :3], gt[:, :3]) and np.array_equal(wt[:, 3] & 7, gt[:, 3] & 7)
                    and np.array_equal(wi, gi) and want.start_level == merged.start_level)
            print("PARITY %s %s %s %s world=%d n=%d nodes=%d shard_sizes=%s" % (
                "OK" if same else "MISMATCH", kind, sampling, tiling, world, n, len(wt), [p[3] for p in parts]),
                flush=True)
            ok = ok and same
        del full, mine
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
