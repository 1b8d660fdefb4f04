"""MIN_DISTANCE across shard faces: per-node selected count of the sharded (virtual ranks) run vs the oracle's
sequential greedy.  usage: python tools/md_face_deviation.py [world] [kind] [n] [tiling] [shard_levels]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import schwarzwald_b200 as sw
from schwarzwald_b200 import synth, distributed
from oracle import sworacle

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
kind = sys.argv[2] if len(sys.argv) > 2 else "terrain"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 800_000
tiling = sys.argv[4] if len(sys.argv) > 4 else "ACCURATE"
sl = int(sys.argv[5]) if len(sys.argv) > 5 else None
kw = dict(side_m=1500.0) if kind == "terrain" else dict(side_m=100.0)
xyz = synth.generate(kind, n, 2, device="cpu", **kw).numpy()
bmin, bmax = sw.cubic_bounds(xyz.min(0), xyz.max(0))
spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
cuts = np.linspace(0, len(xyz), world + 1).astype(int)
parts = [torch.from_numpy(xyz[cuts[r]:cuts[r + 1]].copy()).cuda() for r in range(world)]
extra = dict(shard_levels=sl) if sl else {}
results, infos = distributed.tile_with_virtual_ranks(world, parts, "MIN_DISTANCE", tiling, bmin, bmax, spacing,
                                                     max_points_per_node=3000, concurrency=4, **extra)
got = distributed.merge_results(results)
p = sworacle.make_params("MIN_DISTANCE", tiling, spacing, bmin, bmax, max_points_per_node=3000, concurrency=4)
want = sworacle.Oracle("port").tile(p, xyz)
wc = {(int(x["levels"]), int(x["index"])): int(x["count"]) for x in want.nodes}
shard_levels = infos[0]["shard_levels"]
print("shard_levels", shard_levels, "nodes", len(got.nodes), len(want.nodes))
per_level = {}
for x in got.nodes:
    lv = int(x["levels"])
    ref = wc.get((lv, int(x["index"])), 0)
    a = per_level.setdefault(lv, [0, 0, 0, 0.0])
    a[0] += int(x["count"]); a[1] += ref; a[2] += 1
    if lv < shard_levels and not (x["flags"] & 3) and ref:
        d = (int(x["count"]) - ref) / ref
        a[3] = max(a[3], abs(d))
        if abs(d) > 0.004:
            print("  node", lv, int(x["index"]), "got", int(x["count"]), "ref", ref, "dev %.4f" % d)
for lv in sorted(per_level):
    a = per_level[lv]
    print("level %d: nodes %d got %d ref %d (%.4f) worst spanning node dev %.4f" % (lv, a[2], a[0], a[1], (a[0] - a[1]) / max(a[1], 1), a[3]))
