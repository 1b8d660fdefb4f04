import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import schwarzwald_b200 as sw
from oracle import sworacle as so
port = so.Oracle("port")
rng = np.random.default_rng(7)
def run_case(n, samp, tiling, maxpts, conc, dup=False):
    xyz = rng.random((n,3))*np.array([1000.,800.,60.]) + np.array([5000.,-300.,12.])
    xyz = np.round(xyz, 3)
    xyz[:5] += 5000
    if dup: xyz[10:40] = xyz[10]
    bmin, bmax = sw.cubic_bounds(xyz[5:].min(0), xyz[5:].max(0))
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    p = so.make_params(samp, tiling, spacing, bmin, bmax, max_points_per_node=maxpts, concurrency=conc)
    t0 = time.time(); ref, clamped = port.tile(p, xyz, return_clamped=True); t1 = time.time()
    g = sw.GpuTiler(samp, tiling, bmin, bmax, spacing, max_points_per_node=maxpts, concurrency=conc)
    x2 = xyz.copy()
    res = g.tile(x2); t2 = time.time()
    keys, order = g.keys(n)
    ok_keys = np.array_equal(keys, ref.keys); ok_order = np.array_equal(order, ref.order)
    ok_clamp = np.array_equal(x2, clamped)
    ta, ia = res.canonical(); tb, ib = ref.canonical()
    ta[:,3] &= 6; tb[:,3] &= 6
    ok_nodes = np.array_equal(ta, tb); ok_ids = np.array_equal(ia, ib)
    print(n, samp, tiling, "S", res.start_level, ref.start_level, "nodes", len(res.nodes), len(ref.nodes), "ids", len(res.ids), len(ref.ids),
          "keys", ok_keys, "order", ok_order, "clamp", ok_clamp, "nodes", ok_nodes, "ids", ok_ids, "cpu %.2fs gpu %.2fs" % (t1-t0, t2-t1), flush=True)
    if not (ok_nodes and ok_ids) and ok_nodes is False:
        k = min(len(ta), len(tb))
        bad = np.nonzero((ta[:k] != tb[:k]).any(1))[0][:5]
        print("  first node diffs", bad, ta[bad].tolist(), tb[bad].tolist())
    g.close()
    return ok_keys and ok_order and ok_nodes and ok_ids and ok_clamp
allok = True
for n, maxpts, conc in ((5000, 100, 2), (300000, 2000, 2), (3000000, 20000, 8)):
    for tiling in ("ACCURATE", "FAST"):
        for samp in ("RANDOM_GRID", "GRID_CENTER", "JITTERED"):
            allok &= run_case(n, samp, tiling, maxpts, conc, dup=(n==300000))
print("ALL OK" if allok else "FAILURES")
