"""ctypes front-end for the CPU oracles — TEST INFRASTRUCTURE ONLY.

Two libraries share one C API shape (see oracle/tiler_oracle.cpp / oracle/ref_driver.cpp):

  Oracle("port")  -> oracle/_build/libsworacle.so  our CPU restatement (prefix ``swo_``)
  Oracle("ref")   -> oracle/_ref/libswref.so       the reference's own hot-path TUs compiled
                                                   verbatim from /root/reference (prefix ``swr_``)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (schwarzwald_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

RANDOM_GRID, GRID_CENTER, MIN_DISTANCE, JITTERED, MIN_DISTANCE_FAST = 0, 1, 2, 3, 4
ACCURATE, FAST = 0, 1
TAKE_ALL_WHEN_COUNT_BELOW_MAX, ALWAYS_ADHERE = 0, 1

SAMPLING_NAMES = {"RANDOM_GRID": 0, "GRID_CENTER": 1, "MIN_DISTANCE": 2, "JITTERED": 3, "MIN_DISTANCE_FAST": 4}
TILING_NAMES = {"ACCURATE": 0, "FAST": 1}


class SwParams(C.Structure):
    _fields_ = [
        ("sampling", C.c_int32),
        ("tiling", C.c_int32),
        ("spacing_at_root", C.c_float),
        ("max_depth", C.c_uint32),
        ("max_points_per_node", C.c_uint64),
        ("bounds_min", C.c_double * 3),
        ("bounds_max", C.c_double * 3),
        ("concurrency", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class SwLasTransform(C.Structure):
    _fields_ = [
        ("scale", C.c_double * 3),
        ("offset", C.c_double * 3),
        ("header_min", C.c_double * 3),
        ("header_max", C.c_double * 3),
        ("center", C.c_double * 3),
        ("shift_to_center", C.c_int32),
        ("reserved", C.c_int32),
    ]


def make_las_transform(scale, offset, header_min, header_max, center=None):
    t = SwLasTransform()
    for a in range(3):
        t.scale[a] = float(scale[a])
        t.offset[a] = float(offset[a])
        t.header_min[a] = float(header_min[a])
        t.header_max[a] = float(header_max[a])
        t.center[a] = float(center[a]) if center is not None else 0.0
    t.shift_to_center = 1 if center is not None else 0
    return t


LAS_HEADER_DTYPE = np.dtype([("offset", "<f8", 3), ("scale", "<f8"), ("max", "<f8", 3), ("reserved", "<f8")])

NODE_DTYPE = np.dtype(
    [("index", "<u8"), ("levels", "<u4"), ("flags", "<u4"), ("first", "<u8"), ("count", "<u8")]
)
assert NODE_DTYPE.itemsize == 32


def make_params(sampling, tiling, spacing_at_root, bounds_min, bounds_max, max_points_per_node=20000,
                max_depth=100, concurrency=8):
    p = SwParams()
    p.sampling = SAMPLING_NAMES[sampling] if isinstance(sampling, str) else int(sampling)
    p.tiling = TILING_NAMES[tiling] if isinstance(tiling, str) else int(tiling)
    p.spacing_at_root = float(np.float32(spacing_at_root))
    p.max_depth = int(max_depth)
    p.max_points_per_node = int(max_points_per_node)
    for a in range(3):
        p.bounds_min[a] = float(bounds_min[a])
        p.bounds_max[a] = float(bounds_max[a])
    p.concurrency = int(concurrency)
    return p


def build(kind: str = "port", quiet: bool = True) -> str:
    """Build the requested oracle library with oracle/Makefile; returns its path."""
    target = {"port": "port", "ref": "ref"}[kind]
    path = os.path.join(HERE, "_build/libsworacle.so" if kind == "port" else "_ref/libswref.so")
    if kind == "ref" and not os.path.isdir("/root/reference/schwarzwald"):
        if os.path.exists(path):
            return path  # prebuilt library travelled with the snapshot
        raise FileNotFoundError("oracle/_ref needs /root/reference (or a prebuilt libswref.so)")
    out = subprocess.run(["make", "-C", HERE, target], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    return path


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref/libswref.so")) or os.path.isdir(
        "/root/reference/schwarzwald")


class TileResult:
    """Nodes of one tiling run: name -> original point ids in the node's stored order."""

    def __init__(self, nodes: np.ndarray, ids: np.ndarray, keys=None, order=None, start_level=-1,
                 duplicate_keys=0):
        self.nodes = nodes
        self.ids = ids
        self.keys = keys
        self.order = order
        self.start_level = start_level
        self.duplicate_keys = duplicate_keys

    @staticmethod
    def node_name(index: int, levels: int) -> str:
        """Potree naming, OctreeNodeIndex::to_string_potree (OctreeNodeIndex.h:545-555)."""
        return "r" + "".join(str((index >> (3 * (levels - 1 - l))) & 7) for l in range(levels))

    def as_dict(self):
        out = {}
        for n in self.nodes:
            name = self.node_name(int(n["index"]), int(n["levels"]))
            assert name not in out, "node persisted twice: " + name
            out[name] = self.ids[int(n["first"]): int(n["first"]) + int(n["count"])]
        return out

    def canonical(self):
        """(levels, index)-sorted node table + concatenated ids: order-independent comparison."""
        order = np.lexsort((self.nodes["index"], self.nodes["levels"]))
        nodes = self.nodes[order]
        chunks = [self.ids[int(n["first"]): int(n["first"]) + int(n["count"])] for n in nodes]
        ids = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
        table = np.stack([nodes["levels"].astype(np.uint64), nodes["index"], nodes["count"],
                          nodes["flags"].astype(np.uint64)], axis=1)
        return table, ids


class Oracle:
    def __init__(self, kind: str = "port"):
        self.kind = kind
        self.prefix = "swo_" if kind == "port" else "swr_"
        self.lib = C.CDLL(build(kind))
        f = self._f
        f("expand_bits_by_3", C.c_uint64, [C.c_uint64])
        f("contract_bits_by_3", C.c_uint64, [C.c_uint64])
        f("index_points", None, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p])
        f("octant_bounds", None, [C.c_uint8, C.c_void_p, C.c_void_p, C.c_void_p])
        f("bounds_from_morton_index", None, [C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p])
        f("required_morton_index_depth", C.c_int32,
          [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_float])
        f("partition_child_octants", None, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p])
        f("sample_points", C.c_int64,
          [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32,
           C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p])
        f("tile", C.c_int, [C.POINTER(SwParams), C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)])
        f("tile_repeat", C.c_int, [C.POINTER(SwParams), C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p,
                                   C.POINTER(C.c_void_p)])
        f("tile_batches", C.c_int, [C.POINTER(SwParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                    C.POINTER(C.c_void_p)])
        f("set_threads", None, [C.c_uint32])
        f("set_reference_sort", None, [C.c_int32])
        f("set_start_level_override", None, [C.c_int32])
        f("tile_seconds", C.c_double, [C.c_void_p])
        f("node_count", C.c_uint64, [C.c_void_p])
        f("point_id_count", C.c_uint64, [C.c_void_p])
        f("start_level", C.c_int32, [C.c_void_p])
        f("duplicate_keys", C.c_uint64, [C.c_void_p])
        f("get_nodes", None, [C.c_void_p, C.c_void_p, C.c_void_p])
        f("get_keys", None, [C.c_void_p, C.c_void_p, C.c_void_p])
        f("las_positions", None, [C.c_void_p, C.c_uint64, C.POINTER(SwLasTransform), C.c_void_p])
        f("payload_pnts", None, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p])
        f("payload_las", None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p])
        f("last_error", C.c_char_p, [C.c_void_p])
        f("destroy", None, [C.c_void_p])

    def _f(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, "_" + name, fn)

    # --- primitives -------------------------------------------------------------------------
    def expand_bits_by_3(self, v):
        return int(self._expand_bits_by_3(int(v)))

    def contract_bits_by_3(self, v):
        return int(self._contract_bits_by_3(int(v)))

    @staticmethod
    def _b(bounds):
        bmin = np.ascontiguousarray(bounds[0], dtype=np.float64)
        bmax = np.ascontiguousarray(bounds[1], dtype=np.float64)
        return bmin, bmax

    def index_points(self, xyz, bounds):
        """index_point<21> with ClampToBounds.  Returns (keys, clamped xyz copy)."""
        xyz = np.array(xyz, dtype=np.float64, order="C", copy=True).reshape(-1, 3)
        keys = np.empty(len(xyz), np.uint64)
        bmin, bmax = self._b(bounds)
        self._index_points(xyz.ctypes.data, len(xyz), bmin.ctypes.data, bmax.ctypes.data, keys.ctypes.data)
        return keys, xyz

    def octant_bounds(self, octant, bounds):
        out = np.empty(6, np.float64)
        bmin, bmax = self._b(bounds)
        self._octant_bounds(int(octant), bmin.ctypes.data, bmax.ctypes.data, out.ctypes.data)
        return out[:3].copy(), out[3:].copy()

    def bounds_from_morton_index(self, key, depth, bounds):
        out = np.empty(6, np.float64)
        bmin, bmax = self._b(bounds)
        self._bounds_from_morton_index(int(key), int(depth), bmin.ctypes.data, bmax.ctypes.data,
                                       out.ctypes.data)
        return out[:3].copy(), out[3:].copy()

    def required_morton_index_depth(self, sampling, node_level, bounds, root_max_spacing):
        bmin, bmax = self._b(bounds)
        s = SAMPLING_NAMES[sampling] if isinstance(sampling, str) else sampling
        return int(self._required_morton_index_depth(s, int(node_level), bmin.ctypes.data,
                                                     bmax.ctypes.data, float(np.float32(root_max_spacing))))

    def partition_child_octants(self, keys, level):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        cuts = np.empty(9, np.uint64)
        self._partition_child_octants(keys.ctypes.data, len(keys), int(level), cuts.ctypes.data)
        return cuts

    def sample_points(self, sampling, xyz, keys, ids, node_key, node_level, bounds, spacing_at_root,
                      behaviour=TAKE_ALL_WHEN_COUNT_BELOW_MAX, max_points_per_node=20000):
        """sample_points() on one node.  Returns (n_selected, keys_out, ids_out) or raises."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        ko = np.empty_like(keys)
        io = np.empty_like(ids)
        bmin, bmax = self._b(bounds)
        s = SAMPLING_NAMES[sampling] if isinstance(sampling, str) else sampling
        r = self._sample_points(s, xyz.ctypes.data, keys.ctypes.data, ids.ctypes.data, len(keys),
                                int(node_key), int(node_level), bmin.ctypes.data, bmax.ctypes.data,
                                float(np.float32(spacing_at_root)), int(behaviour),
                                int(max_points_per_node), ko.ctypes.data, io.ctypes.data)
        if r < 0:
            raise RuntimeError("oracle sample_points failed with code %d" % -r)
        return int(r), ko, io

    # --- LAS input transform and writer payloads ------------------
    def las_positions(self, las_xyz, transform):
        """position_from_las_point + the tiler's shift/float32 transformation -> (n,3) float64."""
        las_xyz = np.ascontiguousarray(las_xyz, dtype=np.int32).reshape(-1, 3)
        out = np.empty((len(las_xyz), 3), np.float64)
        self._las_positions(las_xyz.ctypes.data, len(las_xyz), C.byref(transform), out.ctypes.data)
        return out

    def payload_pnts(self, xyz, ids):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.empty((len(ids), 3), np.float32)
        self._payload_pnts(xyz.ctypes.data, ids.ctypes.data, len(ids), out.ctypes.data)
        return out

    def payload_las(self, xyz, ids, nodes, bounds):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        nodes = np.ascontiguousarray(nodes)
        bmin, bmax = self._b(bounds)
        out = np.zeros((len(ids), 3), np.int32)
        headers = np.zeros(len(nodes), LAS_HEADER_DTYPE)
        self._payload_las(xyz.ctypes.data, ids.ctypes.data, nodes.ctypes.data, len(nodes), bmin.ctypes.data,
                          bmax.ctypes.data, out.ctypes.data, headers.ctypes.data)
        return out, headers

    def set_threads(self, n):
        """Worker threads for tile(): the reference's taskflow tasks (per start node / child subtree, chunked
        indexing) on std::threads; the sort stays one std::sort as in the reference.  Default 1."""
        self._set_threads(int(n))

    def set_reference_sort(self, on=True):
        """Timing runs: sort with std::sort exactly as the reference does (tie order unspecified) instead of the
        std::stable_sort that pins the tie order for the parity runs."""
        self._set_reference_sort(1 if on else 0)

    def set_start_level_override(self, level=-1):
        """FAST: tile with this start level instead of estimating it from the batch (-1 = estimate).  Used by the
        subtree parity checks, which tile a few level-3 subtrees of a large cloud with the whole cloud's level."""
        self._set_start_level_override(int(level))

    # --- whole batch --------------------------------------------------------------------------
    def tile(self, params: SwParams, xyz, return_clamped=False, passes=1, copy=True, want_keys=True):
        """One batch through the whole path.  passes > 1 repeats the run over the same point buffer (timing runs:
        res.pass_seconds holds the time of every pass, the result is the last one's); copy=False lets the oracle
        clamp `xyz` in place instead of working on a copy (saves 24 B per point of host memory)."""
        if copy:
            xyz = np.array(xyz, dtype=np.float64, order="C", copy=True).reshape(-1, 3)
        else:
            assert xyz.dtype == np.float64 and xyz.flags["C_CONTIGUOUS"]
            xyz = xyz.reshape(-1, 3)
        h = C.c_void_p()
        seconds = np.zeros(max(1, int(passes)), np.float64)
        rc = self._tile_repeat(C.byref(params), xyz.ctypes.data, len(xyz), int(passes), seconds.ctypes.data,
                               C.byref(h))
        try:
            if rc != 0:
                err = self._last_error(h).decode()
                raise OracleFailure(rc, err)
            nn = self._node_count(h)
            ni = self._point_id_count(h)
            nodes = np.empty(nn, NODE_DTYPE)
            ids = np.empty(ni, np.uint32)
            keys = order = None
            self._get_nodes(h, nodes.ctypes.data, ids.ctypes.data)
            if want_keys:
                keys = np.empty(len(xyz), np.uint64)
                order = np.empty(len(xyz), np.uint32)
                self._get_keys(h, keys.ctypes.data, order.ctypes.data)
            res = TileResult(nodes, ids, keys, order, int(self._start_level(h)),
                             int(self._duplicate_keys(h)))
            res.seconds = float(self._tile_seconds(h))  # index + sort + tiling inside the library
            res.pass_seconds = seconds
        finally:
            self._destroy(h)
        if return_clamped:
            return res, xyz
        return res


def _tile_batches(self, params, xyz, batch_sizes, return_clamped=False):
    """tiling of several batches (TilingAlgorithmV1 / V3 with cached points and a lossless
    in-memory persistence; FAST fixes its start level on the first batch and reconstructs at the end).  `batch_sizes` splits xyz into consecutive batches; point ids are global."""
    xyz = np.array(xyz, dtype=np.float64, order="C", copy=True).reshape(-1, 3)
    offsets = np.concatenate([[0], np.cumsum(np.asarray(batch_sizes, np.int64))]).astype(np.uint64)
    assert int(offsets[-1]) == len(xyz)
    h = C.c_void_p()
    rc = self._tile_batches(C.byref(params), xyz.ctypes.data, len(xyz), offsets.ctypes.data, len(offsets) - 1,
                            C.byref(h))
    try:
        if rc != 0:
            raise OracleFailure(rc, self._last_error(h).decode())
        nodes = np.empty(self._node_count(h), NODE_DTYPE)
        ids = np.empty(self._point_id_count(h), np.uint32)
        self._get_nodes(h, nodes.ctypes.data, ids.ctypes.data)
        res = TileResult(nodes, ids, start_level=int(self._start_level(h)))
    finally:
        self._destroy(h)
    return (res, xyz) if return_clamped else res


Oracle.tile_batches = _tile_batches


class OracleFailure(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("oracle error %d: %s" % (code, msg))
        self.code = code
