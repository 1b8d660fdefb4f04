"""Host-side mirror of the reference's tiler interfaces on top of the C ABI (include/swgpu.h).

Reference constructs mirrored (paths relative to /root/reference/schwarzwald/core):
  * sampling-strategy names      make_sampling_strategy_from_name, tiling/Sampling.h:774-791
  * TilingStrategy               process/Tiler.cpp:189-198 (ACCURATE = V1, FAST = V3)
  * TilerMetaParameters          process/Tiler.h:64-75
  * build_execution_graph / finalize   tiling/TilingAlgorithms.h:70-116
  * node naming                  OctreeNodeIndex::to_string_potree / to_string_entwine,
                                 datastructures/OctreeNodeIndex.h:533-573
  * cubic bounds, spacing        math/AABB.h:50-61, process/TilerProcess.cpp:598-604,
                                 pointcloud/FileStats.cpp:31-37
The compute happens in libswgpu.so; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native

RANDOM_GRID, GRID_CENTER, MIN_DISTANCE, JITTERED = "RANDOM_GRID", "GRID_CENTER", "MIN_DISTANCE", "JITTERED"
MIN_DISTANCE_FAST = "MIN_DISTANCE_FAST"
ACCURATE, FAST = "ACCURATE", "FAST"
_SAMPLING = {RANDOM_GRID: 0, GRID_CENTER: 1, MIN_DISTANCE: 2, JITTERED: 3, MIN_DISTANCE_FAST: 4}
_TILING = {ACCURATE: 0, FAST: 1}

NODE_TERMINAL = 2
NODE_RECONSTRUCTED = 4
NODE_DEEP = 8

NODE_DTYPE = np.dtype(
    [("index", "<u8"), ("levels", "<u4"), ("flags", "<u4"), ("first", "<u8"), ("count", "<u8")])


# sw_las_node_header, include/sw_types.h: what LASPersistence::persist_points writes into a node's LAS header
LAS_HEADER_DTYPE = np.dtype([("offset", "<f8", 3), ("scale", "<f8"), ("max", "<f8", 3), ("reserved", "<f8")])


def las_transform(scale, offset, header_min, header_max, center=None):
    """sw_las_transform: LAS header scale/offset/bounds (position_from_las_point, io/LASFile.cpp:79-94) and,
    when `center` is given, the shift-to-centre + float32 rounding of the 3DTILES pipeline
    (process/TilerProcess.cpp:552-559; center = cubic_bounds.getCenter())."""
    t = native.SwLasTransform()
    for a in range(3):
        t.scale[a] = float(scale[a])
        t.offset[a] = float(offset[a])
        t.header_min[a] = float(header_min[a])
        t.header_max[a] = float(header_max[a])
        t.center[a] = float(center[a]) if center is not None else 0.0
    t.shift_to_center = 1 if center is not None else 0
    return t


class SwgpuError(RuntimeError):
    """Raised for every non-zero return code; the reference throws std::runtime_error instead."""

    def __init__(self, code, message):
        super().__init__("swgpu error %d: %s" % (code, message))
        self.code = code


def cubic_bounds(bmin, bmax):
    """AABB::makeCubic (math/AABB.h:50-61) with the reference's operation order."""
    bmin = np.asarray(bmin, np.float64)
    bmax = np.asarray(bmax, np.float64)
    extent = bmax - bmin
    half = extent.max() / 2
    centre = bmin + extent / 2
    return centre - half, centre + half


def cubic_bounds_at_origin(bmin, bmax):
    """DatasetMetadata::total_bounds_cubic_at_origin (pointcloud/FileStats.cpp:31-37)."""
    cmin, cmax = cubic_bounds(bmin, bmax)
    centre = cmin + (cmax - cmin) / 2
    return cmin - centre, cmax - centre


def spacing_from_diagonal_fraction(cubic_min, cubic_max, fraction=250.0):
    """(float)(cubic extent length / fraction), process/TilerProcess.cpp:598-604."""
    e = np.asarray(cubic_max, np.float64) - np.asarray(cubic_min, np.float64)
    length = np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
    return np.float32(length / fraction)


def node_name(index, levels, convention="potree"):
    """'r' + one octant digit per level (Potree) or 'D-X-Y-Z' (Entwine)."""
    index, levels = int(index), int(levels)
    octants = [(index >> (3 * (levels - 1 - l))) & 7 for l in range(levels)]
    if convention == "potree":
        return "r" + "".join(str(o) for o in octants)
    if convention == "entwine":
        x = y = z = 0
        for o in octants:
            x = (x << 1) | ((o >> 2) & 1)
            y = (y << 1) | ((o >> 1) & 1)
            z = (z << 1) | (o & 1)
        return "%d-%d-%d-%d" % (levels, x, y, z)
    raise ValueError(convention)


class TileResult:
    """Node table + node-major original point ids (what persist_points would have received)."""

    def __init__(self, nodes, ids, start_level=-1):
        self.nodes = nodes
        self.ids = ids
        self.start_level = start_level

    def as_dict(self):
        return {node_name(n["index"], n["levels"]): self.ids[int(n["first"]): int(n["first"]) + int(n["count"])]
                for n in self.nodes}

    def canonical(self):
        order = np.lexsort((self.nodes["index"], self.nodes["levels"]))
        nodes = self.nodes[order]
        chunks = [self.ids[int(n["first"]): int(n["first"]) + int(n["count"])] for n in nodes]
        ids = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
        table = np.stack([nodes["levels"].astype(np.uint64), nodes["index"], nodes["count"],
                          nodes["flags"].astype(np.uint64)], axis=1)
        return table, ids


class GpuTiler:
    """TilingAlgorithmBase-shaped front end: build_execution_graph(points) ... finalize()."""

    def __init__(self, sampling, tiling, bounds_min, bounds_max, spacing_at_root, max_points_per_node=20000,
                 max_depth=100, concurrency=8, device=0):
        self._lib = native.load_library()
        p = native.SwParams()
        p.sampling = _SAMPLING[sampling]
        p.tiling = _TILING[tiling]
        p.spacing_at_root = float(np.float32(spacing_at_root))
        p.max_depth = int(max_depth)
        p.max_points_per_node = int(max_points_per_node)
        for a in range(3):
            p.bounds_min[a] = float(bounds_min[a])
            p.bounds_max[a] = float(bounds_max[a])
        p.concurrency = int(concurrency)
        self.params = p
        self.sampling, self.tiling = sampling, tiling
        self._h = C.c_void_p()
        rc = self._lib.swgpu_create(C.byref(p), int(device), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise SwgpuError(rc, "swgpu_create failed (no CUDA device, or invalid parameters)")
        self._keepalive = None

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.swgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise SwgpuError(rc, self._lib.swgpu_last_error(self._h).decode())

    # -- configuration ----------------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        self._check(self._lib.swgpu_set_stream(self._h, C.c_void_p(int(cuda_stream_handle))))

    def reserve(self, n):
        self._check(self._lib.swgpu_reserve(self._h, int(n)))

    def set_multi_batch(self, on=True):
        """Every later build_execution_graph() is one batch of a multi-batch run against the device-resident
        node store (TilingAlgorithms.cpp:351-492 with cached points); results carry global point ids."""
        self._check(self._lib.swgpu_set_multi_batch(self._h, 1 if on else 0))

    def enable_timing(self, on=True):
        self._check(self._lib.swgpu_enable_timing(self._h, 1 if on else 0))

    def set_deep_node_policy(self, store_whole=True):
        """Nodes where the reference would re-root (TilingAlgorithms.cpp:444-483): fail with error 6 (default) or
        store all remaining points unsampled, flagged NODE_TERMINAL | NODE_DEEP."""
        self._check(self._lib.swgpu_set_deep_node_policy(self._h, 1 if store_whole else 0))

    def set_sort_mode(self, mode=-1):
        """K2: -1 automatic (top-digit passes + segment finish), 0 eight LSD passes, 1..3 passes over the key bits
        from 8 * mode up + segment finish.  The sorted order is identical in every mode."""
        self._check(self._lib.swgpu_set_sort_mode(self._h, int(mode)))

    # -- the two calls of TilingAlgorithmBase --------------------------------------------------------
    def build_execution_graph(self, points):
        """One batch.  `points`: host numpy (n,3) float64 (clamped in place like index_point does) or
        a CUDA torch tensor / object exposing data_ptr() of n*3 float64."""
        if isinstance(points, np.ndarray):
            if points.dtype != np.float64 or not points.flags["C_CONTIGUOUS"]:
                raise ValueError("positions must be C-contiguous float64 (PointBuffer::positions layout)")
            n = points.size // 3
            self._check(self._lib.swgpu_index_batch(self._h, C.c_void_p(points.ctypes.data), n))
        else:
            n = points.numel() // 3
            self._keepalive = points
            self._check(self._lib.swgpu_index_batch_device(self._h, C.c_void_p(points.data_ptr()), n))
        return n

    def build_execution_graph_las(self, las_xyz, transform):
        """One batch that is still in LAS record form: `las_xyz` host numpy (n,3) int32 (laszip_point X/Y/Z)
        or a CUDA tensor of n*3 int32; `transform` from las_transform().  The positions are computed on
        the device exactly as the reference's reader + point transformation compute them."""
        if isinstance(las_xyz, np.ndarray):
            if las_xyz.dtype != np.int32 or not las_xyz.flags["C_CONTIGUOUS"]:
                raise ValueError("LAS coordinates must be C-contiguous int32")
            n = las_xyz.size // 3
            self._check(self._lib.swgpu_index_batch_las(self._h, C.c_void_p(las_xyz.ctypes.data), n, C.byref(transform)))
        else:
            n = las_xyz.numel() // 3
            self._keepalive = las_xyz
            self._check(self._lib.swgpu_index_batch_las_device(self._h, C.c_void_p(las_xyz.data_ptr()), n,
                                                               C.byref(transform)))
        self._n_last = n
        return n

    def positions(self, n):
        """PointBuffer positions of the last batch (n x 3 float64, original order, after clamping)."""
        out = np.empty((int(n), 3), np.float64)
        self._check(self._lib.swgpu_get_positions(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def finalize(self):
        self._check(self._lib.swgpu_finalize(self._h))

    def tile(self, points):
        self.build_execution_graph(points)
        self.finalize()
        return self.result()

    # -- results ------------------------------------------------------------------------------------
    def result_size(self):
        nn, ni = C.c_uint64(), C.c_uint64()
        self._check(self._lib.swgpu_result_size(self._h, C.byref(nn), C.byref(ni)))
        return nn.value, ni.value

    def start_level(self):
        v = C.c_int32()
        self._check(self._lib.swgpu_get_start_level(self._h, C.byref(v)))
        return v.value

    def clamped_count(self):
        v = C.c_uint64()
        self._check(self._lib.swgpu_get_clamped_count(self._h, C.byref(v)))
        return v.value

    def result(self, ids_out=None, nodes_out=None):
        nn, ni = self.result_size()
        nodes = nodes_out if nodes_out is not None else np.empty(nn, NODE_DTYPE)
        ids = ids_out if ids_out is not None else np.empty(ni, np.uint32)
        self._check(self._lib.swgpu_get_nodes(self._h, C.c_void_p(nodes.ctypes.data), C.c_void_p(ids.ctypes.data)))
        return TileResult(nodes[:nn], ids[:ni], self.start_level())

    # -- writer payloads (node-major, same order as result().ids) -----------------------------------
    def payload_pnts(self, out=None):
        """float32 positions as PNTSWriter's PositionAttribute stores them (io/PNTSWriter.cpp:326-342)."""
        _, ni = self.result_size()
        out = out if out is not None else np.empty((ni, 3), np.float32)
        self._check(self._lib.swgpu_get_payload_pnts(self._h, C.c_void_p(out.ctypes.data)))
        return out[:ni]

    def payload_las(self, out=None):
        """(int32 record coordinates, per-node LAS header values) as LASPersistence::persist_points writes
        them (io/LASPersistence.h:119-131,160-163)."""
        nn, ni = self.result_size()
        out = out if out is not None else np.empty((ni, 3), np.int32)
        headers = np.zeros(nn, LAS_HEADER_DTYPE)
        self._check(self._lib.swgpu_get_payload_las(self._h, C.c_void_p(out.ctypes.data),
                                                    C.c_void_p(headers.ctypes.data)))
        return out[:ni], headers

    def payload_pnts_device(self, out_device_ptr):
        """payload_pnts() into a caller-owned device buffer of n_point_ids x 3 float32."""
        self._check(self._lib.swgpu_get_payload_pnts_device(self._h, C.c_void_p(int(out_device_ptr))))

    def payload_las_device(self, out_device_ptr):
        """payload_las() into a caller-owned device buffer of n_point_ids x 3 int32; returns the node headers."""
        nn, _ = self.result_size()
        headers = np.zeros(nn, LAS_HEADER_DTYPE)
        self._check(self._lib.swgpu_get_payload_las_device(self._h, C.c_void_p(int(out_device_ptr)),
                                                           C.c_void_p(headers.ctypes.data)))
        return headers

    def result_device_ids(self, ids_device_ptr):
        nn, _ = self.result_size()
        nodes = np.empty(nn, NODE_DTYPE)
        self._check(self._lib.swgpu_get_nodes_device_ids(self._h, C.c_void_p(nodes.ctypes.data),
                                                         C.c_void_p(int(ids_device_ptr))))
        return nodes

    def keys(self, n):
        keys = np.empty(n, np.uint64)
        order = np.empty(n, np.uint32)
        self._check(self._lib.swgpu_get_keys(self._h, C.c_void_p(keys.ctypes.data), C.c_void_p(order.ctypes.data)))
        return keys, order

    def stats(self):
        s = native.SwgpuStats()
        self._check(self._lib.swgpu_get_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in native.SwgpuStats._fields_}

    # -- stand-alone primitives (device pointers) ---------------------------------------------------
    def morton_encode_device(self, xyz_ptr, n, keys_ptr):
        self._check(self._lib.swgpu_morton_encode_device(self._h, C.c_void_p(int(xyz_ptr)), int(n),
                                                         C.c_void_p(int(keys_ptr))))

    def sort_keys_device(self, keys_ptr, n, order_ptr):
        self._check(self._lib.swgpu_sort_keys_device(self._h, C.c_void_p(int(keys_ptr)), int(n),
                                                     C.c_void_p(int(order_ptr))))

    def gather_attribute_device(self, src_ptr, width, dst_ptr):
        self._check(self._lib.swgpu_gather_attribute_device(self._h, C.c_void_p(int(src_ptr)), int(width),
                                                            C.c_void_p(int(dst_ptr))))


class MultiGpuTiler:
    """Several GPUs from one process (swgpu_multi_*, include/swgpu.h): the whole batch goes in as one host array,
    like Tiler::build_execution_graph_for_indexing hands it to a TilingAlgorithmBase (process/Tiler.cpp:499-527);
    the library cuts it into slices, shuffles the points over NVLink so that every GPU owns whole Morton-prefix
    subtrees, tiles the shards on one host thread per GPU and merges the node tables."""

    def __init__(self, sampling, tiling, bounds_min, bounds_max, spacing_at_root, devices, max_points_per_node=20000,
                 max_depth=100, concurrency=8):
        self._lib = native.load_library()
        p = native.SwParams()
        p.sampling = _SAMPLING[sampling]
        p.tiling = _TILING[tiling]
        p.spacing_at_root = float(np.float32(spacing_at_root))
        p.max_depth = int(max_depth)
        p.max_points_per_node = int(max_points_per_node)
        for a in range(3):
            p.bounds_min[a] = float(bounds_min[a])
            p.bounds_max[a] = float(bounds_max[a])
        p.concurrency = int(concurrency)
        self.params = p
        self.devices = [int(d) for d in devices]
        dev = (C.c_int * len(self.devices))(*self.devices)
        self._h = C.c_void_p()
        rc = self._lib.swgpu_multi_create(C.byref(p), dev, len(self.devices), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise SwgpuError(rc, "swgpu_multi_create failed (no CUDA device, no peer access, or invalid parameters)")
        self._attr_bytes = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.swgpu_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise SwgpuError(rc, self._lib.swgpu_multi_last_error(self._h).decode())

    def set_min_distance_faces(self, on=True):
        self._check(self._lib.swgpu_multi_set_min_distance_faces(self._h, 1 if on else 0))

    def build_execution_graph(self, points, attributes=None):
        """points: host numpy (n, 3) float64, clamped in place; attributes: optional (n, W) uint8, W in 4/8/12/16."""
        if points.dtype != np.float64 or not points.flags["C_CONTIGUOUS"]:
            raise ValueError("positions must be C-contiguous float64 (PointBuffer::positions layout)")
        n = points.size // 3
        ap, ab = None, 0
        if attributes is not None:
            if attributes.dtype != np.uint8 or not attributes.flags["C_CONTIGUOUS"] or attributes.shape[0] != n:
                raise ValueError("attributes must be a C-contiguous (n, W) uint8 array")
            ap, ab = C.c_void_p(attributes.ctypes.data), int(attributes.shape[1])
        self._attr_bytes = ab
        self._check(self._lib.swgpu_multi_index_batch(self._h, C.c_void_p(points.ctypes.data), n, ap, ab))
        return n

    def finalize(self):
        self._check(self._lib.swgpu_multi_finalize(self._h))

    def info(self):
        s, l, c = C.c_int32(), C.c_uint32(), C.c_uint64()
        pts = np.zeros(len(self.devices), np.uint64)
        self._check(self._lib.swgpu_multi_get_info(self._h, C.byref(s), C.byref(l), C.byref(c), C.c_void_p(pts.ctypes.data)))
        return {"start_level": s.value, "shard_levels": l.value, "clamped": c.value, "shard_points": pts}

    def result(self):
        nn, ni = C.c_uint64(), C.c_uint64()
        self._check(self._lib.swgpu_multi_result_size(self._h, C.byref(nn), C.byref(ni)))
        nodes = np.empty(nn.value, NODE_DTYPE)
        ids = np.empty(ni.value, np.uint32)
        self._check(self._lib.swgpu_multi_get_nodes(self._h, C.c_void_p(nodes.ctypes.data), C.c_void_p(ids.ctypes.data)))
        return TileResult(nodes, ids, self.info()["start_level"])

    def rank_result_with_attributes(self, rank):
        """(TileResult of GPU `rank`'s part, node-major attribute records gathered on that GPU)."""
        nn, ni = C.c_uint64(), C.c_uint64()
        self._check(self._lib.swgpu_multi_rank_result_size(self._h, int(rank), C.byref(nn), C.byref(ni)))
        nodes = np.empty(nn.value, NODE_DTYPE)
        ids = np.empty(ni.value, np.uint32)
        attrs = np.empty((ni.value, self._attr_bytes), np.uint8)
        self._check(self._lib.swgpu_multi_get_rank_attributes(self._h, int(rank), C.c_void_p(nodes.ctypes.data),
                                                              C.c_void_p(ids.ctypes.data), C.c_void_p(attrs.ctypes.data)))
        return TileResult(nodes, ids, self.info()["start_level"]), attrs

    def tile(self, points, attributes=None):
        self.build_execution_graph(points, attributes)
        self.finalize()
        return self.result()
