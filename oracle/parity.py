"""Full-size parity checks of a GPU tiling result against the CPU oracle — TEST INFRASTRUCTURE ONLY.

Used by bench.py (the untimed `parity` leg after the timed region) and tests/test_gpu_fullsize.py.  The GPU
result is handed in as (node table on the host, node-major point ids on the host or on the device); nothing
here is on the product path.

  full_parity      the whole result against a whole oracle run, node by node through digests
                   (schwarzwald_b200/verify.py)
  subtree_parity   for clouds the oracle cannot tile whole in reasonable time: the points of a few Morton-prefix
                   subtrees are tiled by the oracle (FAST: with the start level of the whole cloud) and compared
                   with the part of the GPU result inside those subtrees, ids exact, ancestors included for the
                   grid strategies (a sampling cell never leaves a subtree of <= 6 levels)
  min_spacing_check  MIN_DISTANCE: no two points of a sampled node are closer than the node's spacing
                   (SparseGrid.cpp:116-146, GridCell.cpp:41-58), on the MERGED node of a multi-GPU run
"""
from __future__ import annotations

import time

import numpy as np

from . import sworacle
from schwarzwald_b200 import verify

GRID_STRATEGIES = ("RANDOM_GRID", "GRID_CENTER", "JITTERED")


def _ids_slice(ids, first, count):
    part = ids[first:first + count]
    if isinstance(part, np.ndarray):
        return part
    return part.cpu().numpy().view(np.uint32)  # torch int32 tensor on the device


def full_parity(want, nodes, ids):
    """`want`: oracle TileResult of the same input; nodes / ids: the GPU result (ids numpy or device tensor)."""
    wd = verify.node_digests(want.nodes, want.ids)
    gd = verify.node_digests(nodes, ids) if isinstance(ids, np.ndarray) else verify.node_digests_device(nodes, ids)
    # the oracle reports TAKE_ALL / TERMINAL / RECONSTRUCTED in the low three flag bits, like the library
    wd["flags"] &= 7
    gd["flags"] &= 7
    ok, msg = verify.compare_digests(gd, wd)
    return {"checked": True, "ok": bool(ok), "method": "full result vs oracle, per-node digests", "detail": msg,
            "nodes": int(len(wd)), "ids": int(wd["count"].sum()), "digest": "%016x" % verify.result_digest(gd)}


def subtree_parity(orc, sampling, tiling, spacing, bmin, bmax, concurrency, max_points_per_node, xyz_dev, gids_dev,
                   nodes, ids, start_level, depth=3, budget_points=6_000_000, max_subtree_points=3_000_000,
                   max_subtrees=4):
    """xyz_dev: (n, 3) float64 device tensor of the points this GPU tiled; gids_dev: their global ids (int32 tensor
    holding u32) or None when id == row; nodes / ids: this GPU's result with global ids."""
    import torch
    t0 = time.time()
    n = int(xyz_dev.shape[0])
    if n == 0:
        return {"checked": False, "ok": True, "reason": "empty shard"}
    if sampling.startswith("MIN_DISTANCE") and tiling == "ACCURATE":
        return {"checked": False, "ok": True,
                "reason": "MIN_DISTANCE ACCURATE: a node's input depends on its ancestors' greedy across subtree faces"}
    if tiling == "FAST":
        depth = max(depth, 1)
    prefix = verify.subtree_prefixes_device(xyz_dev, bmin, bmax, depth)
    counts = torch.bincount(prefix, minlength=8 ** depth).cpu().numpy()
    # ACCURATE: the ancestors of the subtree must be sampled (not taken whole) in the subtree-only run as well
    picks = verify.choose_subtrees(counts, max_points_per_node, max_subtree_points, budget_points, max_subtrees)
    if not picks:
        return {"checked": False, "ok": True, "reason": "no subtree of suitable size at depth %d" % depth}
    grid = sampling in GRID_STRATEGIES
    params = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=max_points_per_node,
                                  concurrency=concurrency)
    orc.set_start_level_override(start_level if tiling == "FAST" else -1)
    report = {"checked": True, "ok": True, "method": "subtrees of %d levels vs oracle, ids exact" % depth,
              "subtrees": [], "nodes": 0, "ids": 0}
    try:
        for p in picks:
            rows = torch.nonzero(prefix == p).flatten()
            pts = xyz_dev[rows].cpu().numpy()
            gids = (gids_dev[rows].cpu().numpy().view(np.uint32) if gids_dev is not None
                    else rows.cpu().numpy().astype(np.uint32))
            want = orc.tile(params, pts)
            w_ids = gids[want.ids]  # same layout as want.ids, global ids
            members = np.sort(gids) if grid else None
            # oracle side: every node is inside the subtree or one of its ancestors (grid strategies: the
            # subtree's part of the ancestor; otherwise ancestors are dropped on both sides)
            wn, wi = verify.restrict_to_subtree(want.nodes, w_ids, p, depth, member_ids=members)
            gn, gi = verify.restrict_to_subtree(nodes, _LazyIds(ids), p, depth, member_ids=members)
            same_table = (len(wn) == len(gn) and np.array_equal(wn["levels"], gn["levels"]) and
                          np.array_equal(wn["index"], gn["index"]) and np.array_equal(wn["count"], gn["count"]))
            # flags: ancestors of the subtree are take-all candidates by their GLOBAL count only: compare inside
            inside = gn["levels"] >= depth if same_table else None
            same_flags = same_table and np.array_equal(wn["flags"][inside] & 7, gn["flags"][inside] & 7)
            same_ids = same_table and np.array_equal(wi, gi)
            ok = bool(same_table and same_flags and same_ids)
            report["subtrees"].append({"prefix": int(p), "points": int(len(pts)), "nodes": int(len(gn)),
                                       "ids": int(len(gi)), "ok": ok})
            report["nodes"] += int(len(gn))
            report["ids"] += int(len(gi))
            report["ok"] = report["ok"] and ok
    finally:
        orc.set_start_level_override(-1)
    report["seconds"] = round(time.time() - t0, 2)
    return report


class _LazyIds:
    """ids[first:first+count] -> numpy u32, whether the ids live on the host or on the device."""

    def __init__(self, ids):
        self.ids = ids

    def __getitem__(self, sl):
        part = self.ids[sl]
        if isinstance(part, np.ndarray):
            return part
        return part.cpu().numpy().view(np.uint32)


def too_close_pairs(pts, levels, spacing_at_root):
    """Pairs of `pts` (k, 3) closer than the spacing of a node with `levels` levels, by the reference's own test:
    spacing_at_root / 2^levels as a double, narrowed to float and squared in float by SparseGrid (Sampling.h:446-449,
    SparseGrid.cpp:11-14); a pair conflicts when its squared distance is < that (GridCell.cpp:41-58)."""
    from scipy.spatial import cKDTree
    if len(pts) < 2:
        return 0
    s = np.float32(float(np.float32(spacing_at_root)) / (2.0 ** int(levels)))
    thr = float(np.float32(s * s))
    pairs = cKDTree(pts).query_pairs(float(np.sqrt(thr)) * (1.0 + 1e-9), output_type="ndarray")
    if not len(pairs):
        return 0
    d = pts[pairs[:, 0]] - pts[pairs[:, 1]]
    d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    return int((d2 < thr).sum())


def spanning_node_parts(xyz_of_id, nodes, ids, max_levels, max_points=4_000_000):
    """This rank's parts of the sampled nodes with fewer than `max_levels` levels: {(levels, index): positions}."""
    nodes = np.asarray(nodes)
    out, total = {}, 0
    for k in np.nonzero((nodes["levels"] < max_levels) & ((nodes["flags"] & 3) == 0) & (nodes["count"] > 0))[0]:
        f, c = int(nodes["first"][k]), int(nodes["count"][k])
        if total + c > max_points:
            continue
        out[(int(nodes["levels"][k]), int(nodes["index"][k]))] = xyz_of_id(_ids_slice(ids, f, c))
        total += c
    return out


def merged_min_spacing_check(parts_per_rank, spacing_at_root):
    """MIN_DISTANCE on nodes that span GPUs: the min-spacing invariant on the MERGED node (all ranks' parts)."""
    keys = sorted(set(k for parts in parts_per_rank for k in parts))
    out = {"checked": bool(keys), "ok": True, "method": "min-spacing invariant on nodes merged over all ranks (cKDTree)",
           "nodes": []}
    for key in keys:
        pieces = [parts[key] for parts in parts_per_rank if key in parts]
        pts = np.concatenate(pieces, axis=0)
        bad = too_close_pairs(pts, key[0], spacing_at_root)
        out["nodes"].append({"levels": key[0], "index": key[1], "count": int(len(pts)), "ranks": len(pieces),
                             "too_close_pairs": bad})
        out["ok"] = out["ok"] and bad == 0
    out["n_nodes"] = len(out["nodes"])
    out["points"] = int(sum(n["count"] for n in out["nodes"]))
    out["too_close_pairs"] = int(sum(n["too_close_pairs"] for n in out["nodes"]))
    out["nodes"] = sorted(out["nodes"], key=lambda n: (-n["ranks"], n["levels"], n["index"]))[:6]  # keep the line short
    return out


def min_spacing_check(xyz_of_id, nodes, ids, spacing_at_root, max_nodes=6, max_points=3_000_000):
    """MIN_DISTANCE invariant on whole (merged) nodes: no two stored points of a sampled node closer than the
    node's spacing.  xyz_of_id(ids) -> (k, 3) float64 positions.  Checks the largest sampled nodes first."""
    from scipy.spatial import cKDTree
    nodes = np.asarray(nodes)
    sampled = np.nonzero(((nodes["flags"] & 3) == 0) & (nodes["count"] > 1) & (nodes["count"] <= max_points))[0]
    order = sampled[np.argsort(-nodes["count"][sampled].astype(np.int64), kind="stable")][:max_nodes]
    out = {"checked": True, "ok": True, "method": "min-spacing invariant on merged nodes (cKDTree)", "nodes": []}
    for k in order:
        f, c, lv = int(nodes["first"][k]), int(nodes["count"][k]), int(nodes["levels"][k])
        pts = xyz_of_id(_ids_slice(ids, f, c))
        # spacing of a node with `lv` levels: spacing_at_root / 2^lv as a double, narrowed to float and squared in
        # float by SparseGrid (Sampling.h:446-449, SparseGrid.cpp:11-14); the test is squared distance < that
        s = np.float32(float(np.float32(spacing_at_root)) / (2.0 ** lv))
        thr = float(np.float32(s * s))
        tree = cKDTree(pts)
        pairs = tree.query_pairs(float(np.sqrt(thr)) * (1.0 + 1e-9), output_type="ndarray")
        bad = 0
        if len(pairs):
            d = pts[pairs[:, 0]] - pts[pairs[:, 1]]
            d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
            bad = int((d2 < thr).sum())
        out["nodes"].append({"levels": lv, "index": int(nodes["index"][k]), "count": c, "too_close_pairs": bad})
        out["ok"] = out["ok"] and bad == 0
    return out
