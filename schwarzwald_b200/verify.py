"""Result digests and subtree extraction for full-size parity checks (host-side helper, no oracle import).

A tiling result is a node table plus node-major point ids (include/swgpu.h: swgpu_get_nodes).  At BASELINE
sizes (1e8 .. 2e9 points) results are compared through

  * per-node digests: (levels, index, count, flags, order-sensitive 64-bit hash of the node's ids), computable with
    numpy on the host or with torch on the device, identical values either way; and
  * subtree extraction: the points of a few Morton-prefix subtrees of the cloud are pulled out, the caller tiles
    them with the reference's CPU code (tests/ and bench.py own that step), and `restrict_to_subtree` cuts the matching part out of
    the big result.  For the grid strategies a sampling cell never leaves a subtree of <= 6 levels,
    so the part of EVERY node (ancestors included) that lies inside the subtree only depends on the subtree's own
    points; for MIN_DISTANCE this holds for the nodes inside the subtree of a FAST run.

Nothing here touches the oracle: the comparison partner is handed in by the caller.
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1
_K1 = 0x9E3779B97F4A7C15
_K2 = 0xC2B2AE3D27D4EB4F
_K3 = 0xD6E8FEB86659FD93


def _mix_np(ids, pos):
    """64-bit mix of (point id, position inside the node); uint64 arithmetic wraps."""
    with np.errstate(over="ignore"):
        v = (ids.astype(np.uint64) + np.uint64(1)) * np.uint64(_K1) + (pos.astype(np.uint64) + np.uint64(1)) * np.uint64(_K2)
        v ^= v >> np.uint64(31)
        v *= np.uint64(_K3)
        v ^= v >> np.uint64(32)
    return v


def node_digests(nodes, ids):
    """Per-node digests of a result held in host memory.

    Returns a structured array sorted by (levels, index): levels, index, count, flags, digest."""
    nodes = np.asarray(nodes)
    ids = np.asarray(ids)
    out = np.zeros(len(nodes), dtype=[("levels", "<u4"), ("index", "<u8"), ("count", "<u8"), ("flags", "<u4"),
                                      ("digest", "<u8")])
    first = nodes["first"].astype(np.int64)
    count = nodes["count"].astype(np.int64)
    dig = np.zeros(len(nodes), np.uint64)
    if len(nodes) and len(ids):
        # position of every id inside its node
        order = np.lexsort((count, first))  # empty nodes share their successor's offset: they come first
        f_sorted, c_sorted = first[order], count[order]
        assert np.array_equal(f_sorted[1:], (f_sorted + c_sorted)[:-1]) and f_sorted[0] == 0, \
            "node table does not tile the id array"
        chunk = 1 << 26
        for s in range(0, len(ids), chunk):
            e = min(len(ids), s + chunk)
            # nodes overlapping [s, e)
            k0 = int(np.searchsorted(f_sorted, s, side="right")) - 1
            k1 = int(np.searchsorted(f_sorted, e, side="left"))
            seg_first = f_sorted[k0:k1]
            glob = np.arange(s, e, dtype=np.int64)
            owner = np.searchsorted(seg_first, glob, side="right") - 1
            pos = glob - seg_first[owner]
            v = _mix_np(ids[s:e], pos)
            starts = np.maximum(seg_first, s) - s
            nz = c_sorted[k0:k1] > 0
            with np.errstate(over="ignore"):
                sums = np.add.reduceat(v, starts[nz]) if nz.any() else np.zeros(0, np.uint64)
                part = np.zeros(k1 - k0, np.uint64)
                part[nz] = sums
                dig[order[k0:k1]] += part
    out["levels"], out["index"], out["count"], out["flags"], out["digest"] = (
        nodes["levels"], nodes["index"], nodes["count"], nodes["flags"], dig)
    return out[np.lexsort((out["index"], out["levels"]))]


def node_digests_device(nodes, ids_device):
    """node_digests() for point ids that live on the device (torch int32 tensor holding u32 values)."""
    import torch
    nodes = np.asarray(nodes)
    out = np.zeros(len(nodes), dtype=[("levels", "<u4"), ("index", "<u8"), ("count", "<u8"), ("flags", "<u4"),
                                      ("digest", "<u8")])
    dev = ids_device.device
    n_ids = int(ids_device.numel())
    first = nodes["first"].astype(np.int64)
    count = nodes["count"].astype(np.int64)
    dig = np.zeros(len(nodes), np.uint64)

    def s64(x):  # python int (mod 2^64) -> signed 64-bit
        x &= _M64
        return x - (1 << 64) if x >= (1 << 63) else x

    if len(nodes) and n_ids:
        order = np.lexsort((count, first))  # empty nodes share their successor's offset: they come first
        f_sorted, c_sorted = first[order], count[order]
        assert np.array_equal(f_sorted[1:], (f_sorted + c_sorted)[:-1]) and f_sorted[0] == 0
        f_dev = torch.from_numpy(f_sorted).to(dev)
        chunk = 1 << 27
        for s in range(0, n_ids, chunk):
            e = min(n_ids, s + chunk)
            k0 = int(np.searchsorted(f_sorted, s, side="right")) - 1
            k1 = int(np.searchsorted(f_sorted, e, side="left"))
            seg_first = f_dev[k0:k1]
            glob = torch.arange(s, e, dtype=torch.int64, device=dev)
            owner = torch.searchsorted(seg_first, glob, right=True) - 1
            pos = glob - seg_first[owner]
            idv = ids_device[s:e].to(torch.int64) & 0xFFFFFFFF
            v = (idv + 1) * s64(_K1) + (pos + 1) * s64(_K2)
            v = v ^ ((v >> 31) & ((1 << 33) - 1))
            v = v * s64(_K3)
            v = v ^ ((v >> 32) & 0xFFFFFFFF)
            sums = torch.zeros(k1 - k0, dtype=torch.int64, device=dev)
            sums.index_add_(0, owner, v)
            part = sums.cpu().numpy().view(np.uint64)
            with np.errstate(over="ignore"):
                dig[order[k0:k1]] += part
            del glob, owner, pos, idv, v, sums
    out["levels"], out["index"], out["count"], out["flags"], out["digest"] = (
        nodes["levels"], nodes["index"], nodes["count"], nodes["flags"], dig)
    return out[np.lexsort((out["index"], out["levels"]))]


def result_digest(digests):
    """One 64-bit value for a whole result (order-independent over nodes, order-sensitive inside a node)."""
    with np.errstate(over="ignore"):
        v = digests["digest"] ^ (digests["index"] * np.uint64(_K2)) ^ (digests["levels"].astype(np.uint64) << np.uint64(56))
        v = v * np.uint64(_K3) + digests["count"] * np.uint64(_K1) + digests["flags"].astype(np.uint64)
        return int(np.bitwise_xor.reduce(v)) if len(v) else 0


def compare_digests(got, want):
    """(ok, message) for two node_digests() tables."""
    if len(got) != len(want):
        return False, "node count differs: %d vs %d" % (len(got), len(want))
    for field in ("levels", "index", "count", "flags", "digest"):
        bad = np.nonzero(got[field] != want[field])[0]
        if len(bad):
            k = int(bad[0])
            return False, "%d nodes differ in %s (first: levels=%d index=%d: %d vs %d)" % (
                len(bad), field, int(want["levels"][k]), int(want["index"][k]), int(got[field][k]), int(want[field][k]))
    return True, "%d nodes, %d ids identical" % (len(want), int(want["count"].sum()))


# ------------------------------------------------------------------------------------------------------
# subtree extraction
# ------------------------------------------------------------------------------------------------------
def subtree_prefixes_device(xyz_device, bounds_min, bounds_max, depth):
    """Leading `depth` (<= 7) octree levels of every point's MortonIndex64, computed with torch on the device with
    the arithmetic of calculate_morton_index<21> (OctreeAlgorithms.h:64-87: (p - min) * scale, truncate, cap) after
    index_point's clamp.  Returns an int64 tensor of octant paths (x<<2 | y<<1 | z per level)."""
    import torch
    assert 1 <= depth <= 7
    n = xyz_device.shape[0]
    out = torch.zeros(n, dtype=torch.int64, device=xyz_device.device)
    cells = []
    for a in range(3):
        lo, hi = float(bounds_min[a]), float(bounds_max[a])
        scale = 2097152.0 / (hi - lo)
        p = xyz_device[:, a].clamp(lo, hi)
        c = ((p - lo) * scale).to(torch.int64).clamp_(max=(1 << 21) - 1) >> (21 - depth)
        cells.append(c)
    for level in range(depth):
        bit = depth - 1 - level
        octant = (((cells[0] >> bit) & 1) << 2) | (((cells[1] >> bit) & 1) << 1) | ((cells[2] >> bit) & 1)
        out = (out << 3) | octant
    return out


def choose_subtrees(prefix_counts, min_points, max_points, budget_points, max_subtrees=6):
    """Deterministic pick of subtree prefixes whose point count lies in (min_points, max_points], spread over
    the occupied prefix range, total <= budget_points."""
    counts = np.asarray(prefix_counts)
    eligible = np.nonzero((counts > min_points) & (counts <= max_points))[0]
    if not len(eligible):
        return []
    picks, total = [], 0
    step = max(1, len(eligible) // max_subtrees)
    for k in range(step // 2, len(eligible), step):
        p = int(eligible[k])
        if total + int(counts[p]) > budget_points and picks:
            continue
        picks.append(p)
        total += int(counts[p])
        if len(picks) >= max_subtrees:
            break
    return picks


def restrict_to_subtree(nodes, ids, prefix, depth, member_ids=None):
    """The part of a result that lies inside the subtree `prefix` (octant path of `depth` levels).

    Nodes inside the subtree (levels >= depth) are taken whole; ancestors (levels < depth) are filtered to the ids
    in `member_ids` (sorted array of the subtree's point ids) and dropped when member_ids is None.  Returns
    (nodes, ids) with `first` renumbered, sorted by (levels, index)."""
    nodes = np.asarray(nodes)
    lv = nodes["levels"].astype(np.int64)
    idx = nodes["index"].astype(np.uint64)
    inside = np.zeros(len(nodes), bool)
    deep = lv >= depth
    inside[deep] = (idx[deep] >> (3 * (lv[deep] - depth)).astype(np.uint64)) == np.uint64(prefix)
    anc = np.zeros(len(nodes), bool)
    if member_ids is not None:
        sh = lv < depth
        anc[sh] = (np.uint64(prefix) >> (3 * (depth - lv[sh])).astype(np.uint64)) == idx[sh]
    keep = np.nonzero(inside | anc)[0]
    keep = keep[np.lexsort((idx[keep], lv[keep]))]
    out_nodes = np.zeros(len(keep), nodes.dtype)
    chunks, first = [], 0
    for j, k in enumerate(keep):
        f, c = int(nodes["first"][k]), int(nodes["count"][k])
        part = np.asarray(ids[f:f + c])
        if anc[k]:
            part = part[np.isin(part, member_ids, assume_unique=False)]
        out_nodes[j] = (nodes["index"][k], nodes["levels"][k], nodes["flags"][k], first, len(part))
        chunks.append(part)
        first += len(part)
    return out_nodes, (np.concatenate(chunks) if chunks else np.zeros(0, np.uint32))
