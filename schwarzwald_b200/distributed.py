"""Multi-GPU tiling: one process (or thread) per GPU, every GPU owns whole Morton-prefix subtrees.

The reference is a single process; its unit of independent work is the octree node — one taskflow
task per start node (tiling/TilingAlgorithms.cpp:1314-1351) and per >= 100 000-point child
(:499-561).  Sharding the points by the leading octree levels of their Morton key keeps that unit
intact, so every GPU runs the unchanged single-GPU pipeline on its subtrees:

  1. local Morton keys                         swgpu_morton_encode_device  (index_point, clamps in place)
  2. coarse (4-level) prefix histogram         swgpu_prefix_histogram_coarse_device + ONE all-gather (16 KB/rank)
  3. splitters + send/recv count matrix        swgpu_choose_splitters on the summed histogram (host)
  4+5. partition AND exchange in one kernel    swgpu_partition_to_peers_device: every point is written straight
       into its destination's receive buffer (peer memory over NVLink, torch symmetric memory); communicators
       without peer mapping: swgpu_partition_device + all_to_all_single (28 B per point)
  6. single-GPU pipeline on the received points; FAST's start level from the global level-5 counts of the
     sorted keys and the take-all decision of nodes above the shard depth both go through the all-reduce
     hook passed to swgpu_set_shard

Results: every rank reports its nodes with GLOBAL point ids.  Nodes with fewer than `shard_levels`
levels span ranks; their parts concatenated in rank order (= Morton order) are the node
(`merge_results`).  RANDOM_GRID / GRID_CENTER / JITTERED are bit-identical to a single-GPU run:
a sampling cell never spans shards (swgpu_max_shard_levels).  MIN_DISTANCE is exact for every node
that lies inside one shard; nodes above the shard depth are sampled per shard (relaxed across the
shard faces) — see DESIGN.md.

The communicator is an object with: rank, world, all_reduce_sum(tensor) in place,
all_to_all_rows(send, send_counts) -> (recv, recv_counts).  `TorchDistComm` wraps
torch.distributed (NCCL on GPUs, gloo in the CPU tests), `ThreadComm` runs several virtual ranks
as threads on ONE GPU (parity tests on a single-GPU box).
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import native
from .tiler import GpuTiler, TileResult, NODE_DTYPE, SwgpuError

COARSE_LEVELS = 4  # octree levels of the pre-exchange histogram (splitters, count matrix)
PREFIX_BINS = native.PREFIX_BINS


# --------------------------------------------------------------------------------------------------
# communicators
# --------------------------------------------------------------------------------------------------
class TorchDistComm:
    """torch.distributed process group: NCCL all-reduce / all-to-all over NVLink on GPUs."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_reduce_sum(self, t):
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)

    def all_gather(self, t):
        """returns a (world, *t.shape) tensor holding every rank's `t`."""
        import torch
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self._dist.all_gather_into_tensor(out, t, group=self.group)
        return out

    def peer_buffers(self, n_bytes, device):
        """Symmetric receive buffer of n_bytes on every rank, mapped into every process of the group
        (torch symmetric memory: CUDA VMM allocations exchanged once, peer access over NVLink).  Returns
        (own tensor of uint8, [device pointer of rank r's buffer as seen from THIS process]).  Collective."""
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty((int(n_bytes),), dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(t, self.group if self.group is not None else self._dist.group.WORLD)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == t.data_ptr()
        return t, ptrs, hdl

    def all_to_all_rows(self, send, send_counts, recv_counts=None):
        """send: tensor whose rows are ordered by destination; returns (recv, recv_counts).
        recv_counts, when the caller already knows them, saves the count exchange and its sync."""
        import torch
        if recv_counts is None:
            sc = torch.tensor(list(send_counts), dtype=torch.int64, device=send.device)
            rc = torch.empty_like(sc)
            self._dist.all_to_all_single(rc, sc, group=self.group)
            recv_counts = [int(x) for x in rc.cpu().tolist()]
        recv_counts = [int(x) for x in recv_counts]
        recv = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        self._dist.all_to_all_single(recv, send, output_split_sizes=recv_counts,
                                     input_split_sizes=[int(x) for x in send_counts], group=self.group)
        return recv, recv_counts


class ThreadComm:
    """Virtual ranks = threads of one process sharing one device (or the CPU).  Test plumbing."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world

    @classmethod
    def create(cls, world):
        shared = cls._Shared(world)
        return [cls(shared, r) for r in range(world)]

    def _sync(self, t):
        if t.is_cuda:
            import torch
            torch.cuda.synchronize(t.device)

    def all_reduce_sum(self, t):
        sh = self.shared
        self._sync(t)
        sh.slots[self.rank] = t
        sh.barrier.wait()
        total = sh.slots[0].clone()
        for r in range(1, self.world):
            total += sh.slots[r]
        self._sync(t)
        sh.barrier.wait()  # everybody has read every input
        t.copy_(total)
        self._sync(t)
        sh.barrier.wait()

    def all_gather(self, t):
        import torch
        sh = self.shared
        self._sync(t)
        sh.slots[self.rank] = t
        sh.barrier.wait()
        out = torch.stack([sh.slots[r] for r in range(self.world)], dim=0)
        self._sync(out)
        sh.barrier.wait()
        return out

    def peer_buffers(self, n_bytes, device):
        """Virtual ranks share one device: every rank's receive buffer is an ordinary tensor whose pointer is
        valid for all of them, which drives swgpu_partition_to_peers_device exactly as real peer mappings do."""
        import torch
        sh = self.shared
        t = torch.empty(int(n_bytes), dtype=torch.uint8, device=device)
        sh.barrier.wait()  # the previous round of slot users is done
        sh.slots[self.rank] = t
        sh.barrier.wait()
        ptrs = [int(sh.slots[r].data_ptr()) for r in range(self.world)]
        sh.barrier.wait()
        return t, ptrs, None

    def all_to_all_rows(self, send, send_counts, recv_counts=None):
        import torch
        sh = self.shared
        self._sync(send)
        sh.slots[self.rank] = (send, list(int(x) for x in send_counts))
        sh.barrier.wait()
        parts, recv_counts = [], []
        for r in range(self.world):
            buf, counts = sh.slots[r]
            off = sum(counts[:self.rank])
            parts.append(buf[off:off + counts[self.rank]])
            recv_counts.append(counts[self.rank])
        recv = torch.cat(parts, dim=0) if parts else send[:0].clone()
        self._sync(recv)
        sh.barrier.wait()
        return recv, recv_counts


# --------------------------------------------------------------------------------------------------
# host-side pieces (pure functions over the library's host entry points; CPU-testable)
# --------------------------------------------------------------------------------------------------
def estimate_start_level(global_bins, concurrency):
    """estimate_start_node_level_in_octree (TilingAlgorithms.cpp:1473-1535) on global prefix counts."""
    lib = native.load_library()
    bins = np.ascontiguousarray(global_bins, dtype=np.uint32)
    assert bins.size == PREFIX_BINS
    out = C.c_int32()
    rc = lib.swgpu_estimate_start_level(C.c_void_p(bins.ctypes.data), int(concurrency), C.byref(out))
    if rc:
        raise SwgpuError(rc, "swgpu_estimate_start_level")
    return out.value


def choose_splitters(global_bins, n_ranks, shard_levels):
    """first_prefix[r]..first_prefix[r+1] = level-5 prefixes owned by rank r."""
    lib = native.load_library()
    bins = np.ascontiguousarray(global_bins, dtype=np.uint32)
    assert bins.size == PREFIX_BINS
    out = np.zeros(n_ranks + 1, np.uint32)
    rc = lib.swgpu_choose_splitters(C.c_void_p(bins.ctypes.data), int(n_ranks), int(shard_levels),
                                    C.c_void_p(out.ctypes.data))
    if rc:
        raise SwgpuError(rc, "swgpu_choose_splitters")
    return out


def merge_results(parts):
    """Concatenates the per-rank results (rank order) into one TileResult: nodes that span ranks
    are joined, ids in rank order = Morton order."""
    order = {}
    for r, res in enumerate(parts):
        for row in res.nodes:
            key = (int(row["levels"]), int(row["index"]))
            order.setdefault(key, []).append((r, int(row["first"]), int(row["count"]), int(row["flags"])))
    nodes = np.zeros(len(order), NODE_DTYPE)
    chunks, first = [], 0
    for i, key in enumerate(sorted(order)):
        cnt, flags = 0, 0
        for r, f, c, fl in order[key]:
            chunks.append(parts[r].ids[f:f + c])
            cnt += c
            flags |= fl
        nodes[i] = (key[1], key[0], flags, first, cnt)
        first += cnt
    ids = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
    start = max((p.start_level for p in parts), default=-1)
    return TileResult(nodes, ids, start)


# --------------------------------------------------------------------------------------------------
# the sharded tiler
# --------------------------------------------------------------------------------------------------
class ShardedTiler:
    """TilingAlgorithmBase-shaped front end for one rank of a multi-GPU run.

    build_execution_graph(xyz) takes this rank's slice of the batch as a CUDA tensor (n_local, 3)
    float64; global point ids are id_base + local index (id_base defaults to the exclusive prefix
    of the per-rank point counts)."""

    def __init__(self, sampling, tiling, bounds_min, bounds_max, spacing_at_root, max_points_per_node=20000,
                 max_depth=100, concurrency=8, device=0, comm=None, shard_levels=None, exchange="auto"):
        import torch
        self._torch = torch
        self.comm = comm if comm is not None else TorchDistComm()
        self.device = torch.device("cuda", device)
        self.tiler = GpuTiler(sampling, tiling, bounds_min, bounds_max, spacing_at_root,
                              max_points_per_node=max_points_per_node, max_depth=max_depth, concurrency=concurrency,
                              device=device)
        self.sampling, self.tiling, self.concurrency = sampling, tiling, int(concurrency)
        lib = self.tiler._lib
        m = C.c_uint32()
        self.tiler._check(lib.swgpu_max_shard_levels(self.tiler._h, C.byref(m)))
        self.max_shard_levels = int(m.value)
        self.shard_levels = min(int(shard_levels), self.max_shard_levels) if shard_levels else self.max_shard_levels
        if self.shard_levels < 1:
            raise ValueError("spacing too coarse to shard: a sampling cell would span GPUs")
        # device scratch owned here (torch tensors): histogram, node-count exchange buffer
        self._bins = torch.zeros(8 ** COARSE_LEVELS, dtype=torch.int32, device=self.device)
        self._hook = native.ALLREDUCE_FN(self._allreduce_hook)  # keep the callback object alive
        self._face_hook = native.ALLGATHERV_FN(self._allgatherv_hook)
        self._face_keep = None
        # MIN_DISTANCE on nodes that span shards: "faces" = per-shard greedy + conflict resolution across the shard
        # faces (min-spacing invariant holds on the merged node); "none" = per-shard greedy only (round-1 behaviour)
        self.min_distance_faces = "faces"
        self._keep = {}
        self._views = {}
        self.last = {}
        self.profile = False  # True: per-phase CUDA-event times in self.last["phase_ms"] (adds a sync)
        # The exchange step.  "peer": ONE kernel partitions the local points and writes them straight into the
        # destinations' receive buffers over NVLink (swgpu_partition_to_peers_device on symmetric memory);
        # "nccl": partition into a send buffer, then all_to_all_single.  "auto" = peer where the communicator
        # can map peer memory (a real multi-process NCCL group), else nccl.
        self.exchange = exchange
        self._peer = None  # dict(cap, xyz, ids, xyz_ptrs, ids_ptrs, handles)
        self._peer_failed = None
        self._tiny = None
        self._stream_handle = 0

    # -- plumbing -----------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        """The stream of the library kernels AND of every torch-side step of the exchange (histogram reset,
        all-gather, barrier all-reduce, all-to-all, the all-reduce hook): they run in one stream order."""
        self.tiler.set_stream(cuda_stream_handle)
        self._stream_handle = int(cuda_stream_handle)

    def _stream_ctx(self, handle=None):
        """torch.cuda.stream context of the handle's stream (0 = the device's default stream)."""
        torch = self._torch
        h = self._stream_handle if handle is None else int(handle or 0)
        if h:
            return torch.cuda.stream(torch.cuda.ExternalStream(h, device=self.device))
        return torch.cuda.stream(torch.cuda.default_stream(self.device))

    def enable_timing(self, on=True):
        self.tiler.enable_timing(on)

    def close(self):
        self.tiler.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _allreduce_hook(self, ctx, ptr, count, stream):
        """swgpu_allreduce_u32_fn: sums `count` u32 device counters over the ranks, in place."""
        try:
            key = (int(ptr), int(count))
            t = self._views.get(key)
            if t is None:  # the library's counter buffer is stable: wrap it once, no copy
                t = self._torch.as_tensor(_DeviceArray(ptr, int(count)), device=self.device)
                self._views[key] = t
            with self._stream_ctx(stream):  # the library's stream: ordered after its counter kernel
                self.comm.all_reduce_sum(t)
            return 0
        except Exception:  # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return 1

    def _allgatherv_hook(self, ctx, send_ptr, send_bytes, recv_ptr_out, recv_bytes_out, stream):
        """swgpu_allgatherv_fn: gathers `send_bytes` bytes of every rank in rank order into a buffer owned here."""
        try:
            torch = self._torch
            n = int(send_bytes)
            with self._stream_ctx(stream):
                mine = torch.tensor([n], dtype=torch.int64, device=self.device)
                sizes = [int(x) for x in self.comm.all_gather(mine).flatten().cpu().tolist()]
                width = max(max(sizes), 8)
                pad = torch.zeros(width, dtype=torch.uint8, device=self.device)
                if n:
                    pad[:n] = torch.as_tensor(_DeviceBytes(send_ptr, n), device=self.device)
                gathered = self.comm.all_gather(pad)  # (world, width)
                out = torch.empty(max(sum(sizes), 8), dtype=torch.uint8, device=self.device)
                off = 0
                for r, sz in enumerate(sizes):
                    if sz:
                        out[off:off + sz] = gathered[r, :sz]
                    off += sz
                self._face_keep = out  # valid until the next call
            recv_ptr_out[0] = out.data_ptr()
            for r, sz in enumerate(sizes):
                recv_bytes_out[r] = sz
            self.last_face_bytes = self.__dict__.get("last_face_bytes", 0) + n
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return 1

    def _set_shard(self, shard_levels, start_level, ids_ptr, first_prefix):
        t, lib = self.tiler, self.tiler._lib
        t._check(lib.swgpu_set_shard(t._h, shard_levels, int(start_level), self._hook, None, C.c_void_p(ids_ptr)))
        if self.min_distance_faces == "faces" and self.sampling.startswith("MIN_DISTANCE"):
            t._check(lib.swgpu_set_shard_faces(t._h, C.c_void_p(first_prefix.ctypes.data), self.comm.world,
                                               self.comm.rank, self._face_hook, None))
        else:
            t._check(lib.swgpu_set_shard_faces(t._h, None, 0, 0, native.ALLGATHERV_FN(), None))

    def _peer_wanted(self):
        if self.exchange == "nccl" or self.comm.world < 2 or not hasattr(self.comm, "peer_buffers"):
            return False
        return self._peer_failed is None

    def _ensure_peer_buffers(self, need_points, attr_bytes=0):
        """(Re)allocates the symmetric receive buffers for at least need_points points per rank.  Every rank
        calls this with the same value (it comes from the all-gathered count matrix)."""
        if self._peer is not None and self._peer["cap"] >= need_points and self._peer["attr_bytes"] >= attr_bytes:
            return True
        try:
            cap = int(need_points * 1.1) + (1 << 16)
            self._peer = None  # release the old mapping first
            xyz, xyz_ptrs, hx = self.comm.peer_buffers(cap * 24, self.device)
            ids, ids_ptrs, hi = self.comm.peer_buffers(cap * 4, self.device)
            att, att_ptrs, ha = (self.comm.peer_buffers(cap * attr_bytes, self.device) if attr_bytes
                                 else (None, None, None))
            self._peer = {"cap": cap, "xyz": xyz, "ids": ids, "xyz_ptrs": xyz_ptrs, "ids_ptrs": ids_ptrs,
                          "attr": att, "attr_ptrs": att_ptrs, "attr_bytes": attr_bytes, "handles": (hx, hi, ha)}
            return True
        except Exception as e:  # no peer mapping on this system: the NCCL exchange does the same job
            if self.exchange == "peer":
                raise
            self._peer_failed = repr(e)
            self._peer = None
            return False

    # -- the two calls of TilingAlgorithmBase --------------------------------------------------------
    def build_execution_graph(self, xyz, id_base=None, attributes=None):
        """attributes: optional (n_local, W) uint8 CUDA tensor, W in {4, 8, 12, 16}: one attribute record per point
        (PointBuffer's attributes packed by the caller) that travels with the point through the exchange;
        afterwards gather_attributes() returns this rank's node-major attribute payload."""
        with self._stream_ctx():
            return self._build_execution_graph(xyz, id_base, attributes)

    def gather_attributes(self, out=None):
        """Node-major attribute records of this rank's result (same order as result().ids)."""
        torch = self._torch
        att = self._keep.get("attr")
        if att is None:
            raise ValueError("build_execution_graph was called without attributes")
        _, ni = self.tiler.result_size()
        w = att.shape[1]
        out = out if out is not None else torch.empty((max(ni, 1), w), dtype=torch.uint8, device=self.device)
        with self._stream_ctx():
            self.tiler.gather_attribute_device(att.data_ptr(), w, out.data_ptr())
        return out[:ni]

    def _build_execution_graph(self, xyz, id_base=None, attributes=None):
        torch = self._torch
        attr_bytes = 0
        if attributes is not None:
            assert attributes.is_cuda and attributes.dtype == torch.uint8 and attributes.is_contiguous()
            attr_bytes = int(attributes.shape[1])
            assert attributes.shape[0] == xyz.numel() // 3 and attr_bytes in (4, 8, 12, 16)
        t, lib, comm = self.tiler, self.tiler._lib, self.comm
        n = xyz.numel() // 3
        assert xyz.is_cuda and xyz.dtype == torch.float64 and xyz.is_contiguous()
        marks = []

        def mark(name):
            if self.profile:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(torch.cuda.current_stream())
                marks.append((name, ev))

        mark("start")
        # 1. keys of the local slice (clamps outliers in place, like index_point)
        keys = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        t.morton_encode_device(xyz.data_ptr(), n, keys.data_ptr())
        mark("encode")
        # 2. coarse (4-level, 4096-bin) prefix histogram of the local slice, accumulated in shared memory; ONE
        #    all-gather of 16 KB per rank gives every rank all histograms: their sum yields the splitters and,
        #    cut at the splitters, the complete send/recv count matrix — no further count exchange.  (The exact
        #    8^6-bin histogram costs one L2 atomic per key, 1.8 ms per 100 M keys; FAST's start level, which
        #    needs those exact counts, is taken from the sorted keys after the exchange instead, step 6.)
        cl = COARSE_LEVELS
        self._bins.zero_()
        t._check(lib.swgpu_prefix_histogram_coarse_device(t._h, C.c_void_p(keys.data_ptr()), n, cl,
                                                          C.c_void_p(self._bins.data_ptr())))
        all_bins = comm.all_gather(self._bins).cpu().numpy().view(np.uint32).astype(np.int64)  # (world, 8^cl)
        mark("histogram+allgather")
        # 3. splitters from the global histogram (host)
        coarse = all_bins.sum(axis=0)
        n_global = int(coarse.sum())
        if n_global >= 2 ** 32:
            raise ValueError("global point ids are 32 bit: at most 2^32 - 1 points per batch")
        # reference behaviour on degenerate batches (TilingAlgorithms.cpp:253-259, Parallel.h:181-186)
        if n_global == 0:
            raise SwgpuError(7, "tile_internal_node: Got zero points to tile @ node r")
        if self.tiling == "FAST" and n_global < self.concurrency:
            raise SwgpuError(9, "Can't scatter a range that has less than 'scatter_factor' elements!")
        start_level = -1  # FAST: estimated by the library on the global level-5 counts (all-reduce hook)
        unit = 8 ** (6 - cl)  # level-5 prefixes per coarse bin
        fine = np.zeros(PREFIX_BINS, np.uint32)
        fine[::unit] = coarse.astype(np.uint32) if n_global < 2 ** 32 else 0
        shard_levels = min(self.shard_levels, cl)
        first_prefix = choose_splitters(fine, comm.world, shard_levels)
        # send/recv count matrix [source, destination]: per-rank histograms cut at the splitters
        cum = np.zeros((comm.world, 8 ** cl + 1), np.int64)
        np.cumsum(all_bins, axis=1, out=cum[:, 1:])
        cuts = (first_prefix.astype(np.int64) // unit)
        count_matrix = np.diff(cum[:, cuts], axis=1)
        counts = count_matrix.sum(axis=1)
        assert int(counts[comm.rank]) == n
        if id_base is None:
            id_base = int(counts[:comm.rank].sum())
        sc = [int(x) for x in count_matrix[comm.rank]]
        rc = [int(x) for x in count_matrix[:, comm.rank]]
        recv_totals = count_matrix.sum(axis=0)
        if self._peer_wanted() and self._ensure_peer_buffers(int(recv_totals.max()), attr_bytes):
            # 4+5 in ONE kernel: stable partition written straight into the destinations' receive buffers
            # (peer memory over NVLink).  The all-gather above ordered this step after every rank's previous
            # use of its receive buffer; the tiny all-reduce below orders every rank's tiling after all writes.
            world = comm.world
            pk = self._peer
            dst_offsets = np.array([int(count_matrix[:comm.rank, r].sum()) for r in range(world)], np.uint64)
            px = (C.c_void_p * world)(*pk["xyz_ptrs"])
            pi = (C.c_void_p * world)(*pk["ids_ptrs"])
            pa = (C.c_void_p * world)(*pk["attr_ptrs"]) if attr_bytes else None
            t._check(lib.swgpu_set_partition_attributes(t._h, C.c_void_p(attributes.data_ptr() if attr_bytes else 0),
                                                        attr_bytes, None, pa))
            t._check(lib.swgpu_partition_to_peers_device(
                t._h, C.c_void_p(keys.data_ptr()), C.c_void_p(xyz.data_ptr()), n, C.c_void_p(first_prefix.ctypes.data),
                world, int(id_base), px, pi, C.c_void_p(dst_offsets.ctypes.data), None))
            del keys
            mark("partition+exchange (peer kernel)")
            if self._tiny is None:
                self._tiny = torch.zeros(1, dtype=torch.int32, device=self.device)
            comm.all_reduce_sum(self._tiny)
            m = int(recv_totals[comm.rank])
            mark("barrier")
            self._set_shard(shard_levels, start_level, pk["ids_ptrs"][comm.rank] if m else 0, first_prefix)
            self._keep = {"xyz": pk["xyz"][:m * 24].view(torch.float64).view(m, 3),
                          "ids": pk["ids"][:m * 4].view(torch.int32),
                          "attr": pk["attr"][:m * attr_bytes].view(m, attr_bytes) if attr_bytes else None}
            t._check(lib.swgpu_index_batch_device(t._h, C.c_void_p(pk["xyz_ptrs"][comm.rank] if m else 0), m))
            mark("tile")
            phases = {}
            if self.profile:
                torch.cuda.synchronize(self.device)
                for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                    phases[name] = e0.elapsed_time(e1)
            self.last = {"phase_ms": phases, "n_local": n, "n_shard": m, "n_global": n_global,
                         "start_level": t.start_level(), "first_prefix": first_prefix, "send_counts": sc, "recv_counts": rc,
                         "shard_levels": shard_levels,
                         "exchange": "peer kernel (swgpu_partition_to_peers_device over symmetric memory)",
                         "bytes_sent_off_gpu": int(sum(c for r, c in enumerate(sc) if r != comm.rank)) * (28 + attr_bytes)}
            return n
        # 4. stable partition into the send buffer
        send_xyz = torch.empty((max(n, 1), 3), dtype=torch.float64, device=self.device)
        send_ids = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        send_counts = np.zeros(comm.world, np.uint64)
        send_attr = torch.empty((max(n, 1), attr_bytes), dtype=torch.uint8, device=self.device) if attr_bytes else None
        t._check(lib.swgpu_set_partition_attributes(t._h, C.c_void_p(attributes.data_ptr() if attr_bytes else 0),
                                                    attr_bytes, C.c_void_p(send_attr.data_ptr() if attr_bytes else 0),
                                                    None))
        t._check(lib.swgpu_partition_device(t._h, C.c_void_p(keys.data_ptr()), C.c_void_p(xyz.data_ptr()), n,
                                            C.c_void_p(first_prefix.ctypes.data), comm.world, int(id_base),
                                            C.c_void_p(send_xyz.data_ptr()), C.c_void_p(send_ids.data_ptr()),
                                            C.c_void_p(send_counts.ctypes.data)))
        del keys
        mark("partition")
        # 5. the exchange: every GPU receives whole subtrees, sources in rank order
        assert sc == [int(x) for x in send_counts], "partition and histogram disagree"
        recv_xyz, recv_counts = comm.all_to_all_rows(send_xyz[:n], sc, rc)
        recv_ids, _ = comm.all_to_all_rows(send_ids[:n], sc, rc)
        recv_attr = comm.all_to_all_rows(send_attr[:n], sc, rc)[0] if attr_bytes else None
        del send_xyz, send_ids, send_attr
        m = int(recv_xyz.shape[0])
        mark("all_to_all")
        # 6. the single-GPU pipeline on the shard
        self._set_shard(shard_levels, start_level, recv_ids.data_ptr() if m else 0, first_prefix)
        self._keep = {"xyz": recv_xyz, "ids": recv_ids, "attr": recv_attr}
        t._check(lib.swgpu_index_batch_device(t._h, C.c_void_p(recv_xyz.data_ptr() if m else 0), m))
        mark("tile")
        phases = {}
        if self.profile:
            torch.cuda.synchronize(self.device)
            for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                phases[name] = e0.elapsed_time(e1)
        self.last = {"phase_ms": phases, "n_local": n, "n_shard": m, "n_global": n_global, "start_level": t.start_level(),
                     "shard_levels": shard_levels, "first_prefix": first_prefix, "send_counts": sc, "recv_counts": recv_counts,
                     "exchange": "nccl all_to_all" + (" (peer mapping unavailable: %s)" % self._peer_failed
                                                      if self._peer_failed else ""),
                     "bytes_sent_off_gpu": int(sum(c for r, c in enumerate(sc) if r != comm.rank)) * (28 + attr_bytes)}
        return n

    def finalize(self):
        self.tiler.finalize()

    # -- results ------------------------------------------------------------------------------------
    def result(self, **kw):
        """This rank's nodes with global point ids (spanning nodes: this rank's part)."""
        return self.tiler.result(**kw)

    def result_size(self):
        return self.tiler.result_size()

    def result_device_ids(self, ids_device_ptr):
        return self.tiler.result_device_ids(ids_device_ptr)

    def stats(self):
        return self.tiler.stats()

    def start_level(self):
        return self.tiler.start_level()

    def shard_positions(self):
        """The received (shard-resident) positions and their global ids."""
        return self._keep.get("xyz"), self._keep.get("ids")


class _DeviceBytes:
    """__cuda_array_interface__ view of `count` bytes at a raw device pointer (no copy)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class _DeviceArray:
    """__cuda_array_interface__ view of `count` int32 at a raw device pointer (no copy)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def tile_with_virtual_ranks(world, xyz_parts, sampling, tiling, bounds_min, bounds_max, spacing_at_root, device=0,
                            attr_parts=None, **kw):
    """Runs `world` virtual ranks as threads on one GPU (ThreadComm) and returns the per-rank
    TileResults plus the per-rank bookkeeping.  Used by the single-GPU parity tests."""
    import torch
    comms = ThreadComm.create(world)
    results, infos, errors = [None] * world, [None] * world, [None] * world

    def work(r):
        try:
            torch.cuda.set_device(device)
            with ShardedTiler(sampling, tiling, bounds_min, bounds_max, spacing_at_root, device=device,
                              comm=comms[r], **kw) as st:
                st.build_execution_graph(xyz_parts[r], attributes=attr_parts[r] if attr_parts else None)
                st.finalize()
                results[r] = st.result()
                infos[r] = dict(st.last)
                if attr_parts:
                    infos[r]["attributes"] = st.gather_attributes().cpu().numpy()
        except BaseException as e:  # noqa: BLE001 - reported to the caller
            errors[r] = e
            comms[r].shared.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e
    return results, infos
