/*
 * swgpu.h — C ABI of libswgpu, the B200 (sm_100a) implementation of Schwarzwald's tiler compute
 * core: Morton indexing, the node-grouping sort and per-node LOD sampling for ACCURATE and FAST
 * tiling.  Plain pointers and sizes only; no CUDA, torch or C++ types cross this boundary.
 *
 * The reference has no FFI; its seam is the C++ interface TilingAlgorithmBase
 * (schwarzwald/core/tiling/TilingAlgorithms.h:70-116).  Each entry point below names the
 * reference construct it stands in for (paths relative to /root/reference/schwarzwald/core).
 * A drop-in `TilingAlgorithmGPU : TilingAlgorithmBase` built on these calls is shown in
 * INTEGRATION.md and shipped as schwarzwald_b200/host/TilingAlgorithmGPU.h.
 *
 * Conventions: every function returns SW_OK (0) or an SW_ERR_* code (include/sw_types.h); the
 * message of the last failure is available from swgpu_last_error().  The caller owns all host
 * buffers, the library owns all device memory.  A handle is used from one host thread at a time
 * (the reference runs one indexing task per batch, process/Tiler.cpp:499-527).  There is no CPU
 * fallback: without a CUDA device every compute call fails with SW_ERR_CUDA.
 */
#ifndef SWGPU_H
#define SWGPU_H

#include "sw_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swgpu_tiler* swgpu_handle;

/* TilingAlgorithmV1 / TilingAlgorithmV3 constructors (tiling/TilingAlgorithms.cpp:567-575,
 * 1195-1205) + Tiler::Tiler's choice between them (process/Tiler.cpp:189-198).  `device` is the
 * CUDA device ordinal. */
int swgpu_create(const sw_params* params, int device, swgpu_handle* out);
void swgpu_destroy(swgpu_handle h);
const char* swgpu_last_error(swgpu_handle h);

/* Work is enqueued on this CUDA stream (a cudaStream_t passed as void*; NULL = default stream). */
int swgpu_set_stream(swgpu_handle h, void* cuda_stream);

/* Pre-allocates device memory for batches of up to n points (optional; batches grow on demand). */
int swgpu_reserve(swgpu_handle h, uint64_t n);

/*
 * build_execution_graph() for one batch (TilingAlgorithms.cpp:577-626 for ACCURATE, 1250-1360 for
 * FAST's first iteration): index every point, sort, run the per-node sampling.
 *   xyz  AoS n x 3 doubles = PointBuffer::positions() (datastructures/PointBuffer.h:291).
 *        Points outside the bounds are clamped IN PLACE, as index_point does
 *        (tiling/OctreeAlgorithms.h:156-170).
 * swgpu_index_batch takes a HOST buffer (copied to the device, clamped values copied back);
 * swgpu_index_batch_device takes a DEVICE pointer that stays owned by the caller and must stay
 * valid until the results have been fetched.  Device pointers handed to the indexing entry points
 * (positions, LAS records, keys) must be 16-byte aligned (any cudaMalloc / whole-tensor pointer is);
 * others are refused with SW_ERR_INVALID_ARGUMENT.
 */
int swgpu_index_batch(swgpu_handle h, double* xyz_host, uint64_t n);
int swgpu_index_batch_device(swgpu_handle h, double* xyz_device, uint64_t n);

/* TilingAlgorithmBase::finalize(bounds): FAST re-samples the skipped upper levels from their
 * children (reconstruct_left_out_nodes, TilingAlgorithms.cpp:1717-1784); no-op for ACCURATE. */
int swgpu_finalize(swgpu_handle h);

/* Multi-batch mode (the reference's default: internal_cache_size = 10 M points per batch,
 * executable/main.cpp:233-236).  After swgpu_set_multi_batch(h, 1) every swgpu_index_batch* call is one
 * build_execution_graph of TilingAlgorithmV1 / V3 against what the earlier batches stored: a node that is
 * reached again gets its stored points back, re-keyed relative to the node bounds (read_pnts_from_disk,
 * TilingAlgorithms.cpp:50-109), merged with the incoming ones (incoming first on equal keys, tiling/Node.cpp:3-34)
 * and is sampled again with AlwaysAdhereToMinSpacing (:272-275); FAST fixes its start level on the first batch
 * (:1250-1360), cuts later batches at it (:1362-1453) and reconstructs the upper levels in swgpu_finalize from
 * everything stored (:1717-1784).  The store (all positions, clamped, and the node tables) lives in device
 * memory; point ids are global: the i-th point of a batch has id (points of all earlier batches) + i, all
 * batches together must stay below 2^32 points.  swgpu_result_size / swgpu_get_nodes* then return the final
 * content of every node, ordered by (levels, index).  Switching the mode (on or off) empties the store.
 * Not combinable with swgpu_set_shard. */
int swgpu_set_multi_batch(swgpu_handle h, int enable);

/* Nodes below the capacity of a MortonIndex64.  A node at level >= 14 (grid strategies, JITTERED; level 20 for
 * MIN_DISTANCE never gets there) whose sampling grid would need more than 21 key levels is re-indexed by the
 * reference with the node as the new root (TilingAlgorithms.cpp:444-483).  That happens when more than
 * max_points_per_node points share a cell of 1 / 2^15 of the extent, or - in multi-batch runs, where a revisited
 * node is always sampled (:272-275) - as soon as two points of different batches cannot be told apart above that
 * level (duplicates, millimetre-quantised LiDAR).  The library does not reproduce the re-root:
 *   policy 0 (default)  the batch fails with SW_ERR_DEEP_REROOT, like the oracle;
 *   policy 1            such a node stores all its remaining points unsampled and is flagged
 *                       SW_NODE_TERMINAL | SW_NODE_DEEP - the same content the reference's re-rooted sampling
 *                       produces whenever it can separate the points (each in its own cell of the finer grid),
 *                       a stated deviation otherwise. */
int swgpu_set_deep_node_policy(swgpu_handle h, int policy);

/* K2, the sort that replaces std::sort over IndexedPoint64 (TilingAlgorithms.cpp:600-604,1289-1292).  The result is
 * the same in every mode (key order, ties by original point index); the modes differ in how often the pairs move:
 *   -1 (default)  automatic: the first batch of a handle is sorted by the eight LSD passes and its sorted keys are
 *                 probed for the lengths of the runs of equal top 40 / 48 key bits; while those are short (terrain,
 *                 uniform and volume-filling clouds) later batches run onesweep passes over the top 40 or 48 bits
 *                 only and order the runs in place with one kernel; clustered clouds keep the eight passes
 *    0            eight LSD passes over all 63 bits
 *    1, 2, 3      passes over the key bits from 8 * mode up, then the segment finish */
int swgpu_set_sort_mode(swgpu_handle h, int mode);

/* The hand-off that replaces the per-node persist_points() calls (io/PointsPersistence.h:23-31):
 * a node table plus one node-major array of ORIGINAL point indices, Morton-ordered inside each
 * node.  The adapter turns row i into persist_points(refs[first..first+count), bounds, name). */
int swgpu_result_size(swgpu_handle h, uint64_t* n_nodes, uint64_t* n_point_ids);
int swgpu_get_nodes(swgpu_handle h, sw_node* nodes, uint32_t* point_ids);
/* Same hand-off without leaving the device: point ids are written to a caller-owned device buffer
 * of n_point_ids u32 (node table still goes to the host). */
int swgpu_get_nodes_device_ids(swgpu_handle h, sw_node* nodes, uint32_t* point_ids_device);

/* FAST: _level_of_start_nodes (TilingAlgorithms.cpp:1294-1295); -1 for ACCURATE. */
int swgpu_get_start_level(swgpu_handle h, int32_t* level);
/* number of points index_point clamped in the last batch */
int swgpu_get_clamped_count(swgpu_handle h, uint64_t* n);

/* Test hooks: the sorted Morton keys and the sort permutation (original index of each sorted
 * position) of the last batch.  Either pointer may be NULL. */
int swgpu_get_keys(swgpu_handle h, uint64_t* keys, uint32_t* order);

/* Optional attribute permutation (the gather persist_points performs through PointReference):
 * gathers `n_records` records of `width` bytes (1,2,3,4,8,12 or 24) from a device array indexed
 * by original point index into node-major order on the device.  src/dst are device pointers. */
int swgpu_gather_attribute_device(swgpu_handle h, const void* src_device, uint32_t width, void* dst_device);

/*
 * ---- LAS coordinates in ------------------------------------------------------
 * build_execution_graph() for a batch that is still in LAS record form: `las_xyz` holds n x 3 int32
 * (laszip_point::X, Y, Z).  The device computes the PointBuffer positions exactly as the reference's
 * reader and point transformation do (sw_las_transform, include/sw_types.h: io/LASFile.cpp:79-94,
 * process/TilerProcess.cpp:552-559) and indexes them in the same kernel: the host pass over every point
 * and half of the host-to-device bytes disappear.  swgpu_get_positions returns those positions
 * (n x 3 doubles, original point order, after index_point's clamping) for callers that still need them
 * on the host; it works after any swgpu_index_batch* call.
 */
int swgpu_index_batch_las(swgpu_handle h, const int32_t* las_xyz_host, uint64_t n, const sw_las_transform* t);
int swgpu_index_batch_las_device(swgpu_handle h, const int32_t* las_xyz_device, uint64_t n, const sw_las_transform* t);
int swgpu_get_positions(swgpu_handle h, double* xyz_host);

/*
 * ---- Writer payloads ---------------------------------------------------------
 * The position bytes the reference's writers store, for ALL nodes at once in the node-major order of
 * swgpu_get_nodes (row i of the node table owns payload records [first, first + count)):
 *   pnts  n_point_ids x 3 float32 = attributes::PositionAttribute (io/PNTSWriter.cpp:326-342)
 *   las   n_point_ids x 3 int32 record coordinates + one sw_las_node_header per node table row
 *         (io/LASPersistence.h:119-131,160-163; io/LASPersistence.cpp:17-28; quantisation of LASzip's
 *         laszip_set_coordinates)
 * The _device variants write to caller-owned device buffers (headers still go to the host).
 */
int swgpu_get_payload_pnts(swgpu_handle h, float* xyz_f32_host);
int swgpu_get_payload_pnts_device(swgpu_handle h, float* xyz_f32_device);
int swgpu_get_payload_las(swgpu_handle h, int32_t* xyz_i32_host, sw_las_node_header* headers_host);
int swgpu_get_payload_las_device(swgpu_handle h, int32_t* xyz_i32_device, sw_las_node_header* headers_host);

/* Stand-alone primitives (used by the parity tests and the multi-GPU shuffle). */
/* index_point<21> over a device batch: keys_device receives n u64; xyz is clamped in place. */
int swgpu_morton_encode_device(swgpu_handle h, double* xyz_device, uint64_t n, uint64_t* keys_device);
/* Stable sort of (key, original index) on the device; keys sorted in place, order_device gets n u32. */
int swgpu_sort_keys_device(swgpu_handle h, uint64_t* keys_device, uint64_t n, uint32_t* order_device);

/*
 * ---- Multi-GPU: one handle per GPU / process, every GPU tiles whole Morton-prefix subtrees -------
 * The reference is a single process; its unit of independent work is the octree node (one taskflow
 * task per start node, tiling/TilingAlgorithms.cpp:1314-1351, :499-561).  These entry points let a
 * host driver (schwarzwald_b200/distributed.py over torch.distributed/NCCL, or a C++ driver over
 * ncclSend/ncclRecv) shard the points by the leading octree levels of their Morton key:
 *   1. swgpu_morton_encode_device            local keys (index_point, clamps in place)
 *   2. swgpu_prefix_histogram_coarse_device  local counts of the leading 4 octree levels; the caller all-gathers
 *                                            them (16 KB per rank): sum = global histogram, cut at the splitters =
 *                                            the complete send/receive count matrix
 *      (or swgpu_prefix_histogram_device     the exact 8^6-bin variant, one L2 atomic per key)
 *   3. swgpu_choose_splitters                contiguous prefix ranges of (nearly) equal point count
 *   4. swgpu_partition_to_peers_device       stable partition written straight into the destinations' receive
 *                                            buffers (peer memory over NVLink): partition AND exchange in one kernel
 *      (or swgpu_partition_device + the caller's all-to-all, where peer memory cannot be mapped)
 *   5. swgpu_set_shard + swgpu_index_batch_device on the received points, swgpu_finalize; FAST's start level
 *      comes from the global level-5 counts of the sorted keys through the all-reduce hook (start_level = -1)
 *      or from swgpu_estimate_start_level on an exact global histogram
 * Nodes with fewer than `shard_levels` levels span GPUs: every GPU reports its part of such a node
 * (same index/levels, ids in Morton order); the parts concatenated in rank order are the node.
 * Their take-all decision (Sampling.h:201-208) uses the global point count, summed through the
 * caller's all-reduce hook.  Grid strategies stay bit-exact because a sampling cell never spans
 * shards as long as shard_levels <= swgpu_max_shard_levels().
 */
#define SWGPU_MAX_RANKS 16
#define SWGPU_PREFIX_BINS 262144

/* In-place SUM all-reduce of `count` u32 device counters over all ranks, enqueued on (or ordered
 * with) `cuda_stream`.  Returns 0 on success.  Called once per sweep level below shard_levels by
 * every rank, in the same order on every rank. */
typedef int (*swgpu_allreduce_u32_fn)(void* ctx, uint32_t* device_counters, uint64_t count, void* cuda_stream);

int swgpu_prefix_histogram_device(swgpu_handle h, const uint64_t* keys_device, uint64_t n, uint32_t* bins_device);
/* The same for the leading `levels` (1..4) octree levels: 8^levels bins, accumulated in shared memory (one read
 * of the keys, no L2 atomic per key).  Enough for balanced splitters; with start_level = -1 in swgpu_set_shard
 * the exact level-5 counts FAST's start level needs are taken from the sorted keys after the exchange and
 * summed through the all-reduce hook. */
int swgpu_prefix_histogram_coarse_device(swgpu_handle h, const uint64_t* keys_device, uint64_t n, uint32_t levels,
                                         uint32_t* bins_device);
/* estimate_start_node_level_in_octree (TilingAlgorithms.cpp:1473-1535) on GLOBAL level-5 prefix
 * counts (host array of SWGPU_PREFIX_BINS).  Pure host function. */
int swgpu_estimate_start_level(const uint32_t* bins_host, uint32_t concurrency, int32_t* level);
/* first_prefix[r] .. first_prefix[r+1] = level-5 prefixes owned by rank r (n_ranks + 1 entries),
 * boundaries snapped to multiples of 8^(6 - shard_levels).  Pure host function. */
int swgpu_choose_splitters(const uint32_t* bins_host, uint32_t n_ranks, uint32_t shard_levels, uint32_t* first_prefix);
/* Deepest shard prefix for which every sampling cell of this handle's strategy lies inside one
 * shard (<= 6). */
int swgpu_max_shard_levels(swgpu_handle h, uint32_t* shard_levels);
/* Stable partition of the local points by destination rank.  out_xyz_device (n x 3 doubles) and
 * out_id_device (n u32 = id_base + local index) are ordered by destination; send_counts_host
 * receives n_ranks counts. */
int swgpu_partition_device(swgpu_handle h, const uint64_t* keys_device, const double* xyz_device, uint64_t n,
                           const uint32_t* first_prefix, uint32_t n_ranks, uint32_t id_base, double* out_xyz_device,
                           uint32_t* out_id_device, uint64_t* send_counts_host);
/* Partition + exchange in ONE kernel over peer memory: the same stable partition, written straight
 * into the destinations' receive buffers.  peer_xyz_device[r] / peer_ids_device[r] are device
 * pointers to rank r's receive buffers as mapped into THIS process (CUDA IPC / VMM peer mappings over
 * NVLink; the local buffers for r == own rank); dst_offsets[r] is the first point of this source's
 * block inside rank r's buffer (= points rank r receives from lower ranks; known from the all-gathered
 * prefix histograms).  The caller orders the kernel against the peers (a barrier before, so that the
 * receive buffers are free; one after, before anybody reads).  send_counts_host may be NULL. */
int swgpu_partition_to_peers_device(swgpu_handle h, const uint64_t* keys_device, const double* xyz_device, uint64_t n,
                                    const uint32_t* first_prefix, uint32_t n_ranks, uint32_t id_base,
                                    void* const* peer_xyz_device, void* const* peer_ids_device,
                                    const uint64_t* dst_offsets, uint64_t* send_counts_host);
/* Attributes through the exchange (PointBuffer's per-point attributes, core/datastructures/PointBuffer.h:291-304:
 * 14 bytes per point for LAS point format 2, packed by the caller into one record of 4, 8, 12 or 16 bytes per
 * point).  The next swgpu_partition_device / swgpu_partition_to_peers_device call moves record i with point i:
 * into out_attr_device (send buffer, same order as out_xyz_device) or into the ranks' receive buffers
 * peer_attr_device[r] (same offsets as the positions; the array must stay valid until that call returns).
 * After the exchange every rank holds the attributes of the points it tiles, indexed like its positions, so
 * swgpu_gather_attribute_device produces the node-major attribute payload locally.  attr_device = NULL
 * switches it off. */
int swgpu_set_partition_attributes(swgpu_handle h, const void* attr_device, uint32_t attr_bytes, void* out_attr_device,
                                   void* const* peer_attr_device);
/* Marks the handle as tiling one shard.  start_level: FAST's global start level; -1 = the library estimates
 * it itself: on the GLOBAL level-5 counts (one 1 MB all-reduce through the hook, collective over all ranks)
 * when a hook is given, on the local points otherwise; global_ids_device: id of every received point (returned by swgpu_get_nodes
 * instead of local indices; may be NULL).  shard_levels = 0 switches sharding off. */
int swgpu_set_shard(swgpu_handle h, uint32_t shard_levels, int32_t start_level, swgpu_allreduce_u32_fn allreduce,
                    void* allreduce_ctx, const uint32_t* global_ids_device);

/* MIN_DISTANCE / MIN_DISTANCE_FAST on nodes that span shards (fewer than shard_levels levels; FAST's reconstructed
 * levels as well).  PoissonDiskSampling (core/tiling/Sampling.h:421-471, core/datastructures/SparseGrid.cpp:116-146)
 * is ONE greedy over all points of a node in Morton order; running it in rank order would serialise the GPUs.
 * With this hook every shard samples its part of the node on its own (exact inside the part), the accepted points
 * that have another shard within the spacing are all-gathered (32 bytes each: key + position), and a point that is
 * closer than the spacing to an accepted point of a LOWER rank (earlier in Morton order) is rejected and moves on
 * to the child nodes.  The merged node then keeps the reference's invariant — no two stored points closer than
 * the spacing — and may hold slightly fewer points than the sequential greedy (north_star's relaxed mode; the
 * stated bound of 1 % per node is checked by tests/test_gpu_sharded.py).  Nodes inside one shard stay exact.
 * Without the hook, spanning nodes are sampled per shard with no exchange (the invariant can then fail across
 * shard faces).  first_prefix / n_ranks: the splitters of swgpu_choose_splitters; rank: this shard.
 * The hook gathers `send_bytes` bytes of every rank, in rank order, into a device buffer it owns and that stays
 * valid until its next call: *recv_device = that buffer, recv_bytes[r] = bytes of rank r; all work on
 * cuda_stream.  Collective: the library calls it once per sampled spanning level on every rank. */
typedef int (*swgpu_allgatherv_fn)(void* ctx, const void* send_device, uint64_t send_bytes, void** recv_device,
                                   uint64_t* recv_bytes, void* cuda_stream);
int swgpu_set_shard_faces(swgpu_handle h, const uint32_t* first_prefix, uint32_t n_ranks, uint32_t rank,
                          swgpu_allgatherv_fn allgatherv, void* ctx);

/* ---- several GPUs from ONE host process ---------------------------------------------------------------------
 * The reference is one process (core/process/Tiler.cpp:189-198, 499-527); these calls let that process tile a
 * batch on n GPUs of the box: the batch is cut into n slices, each GPU indexes its slice, the points are shuffled
 * over NVLink (peer access) so that every GPU owns whole Morton-prefix subtrees, every GPU runs the single-GPU
 * pipeline on its shard (one host thread per GPU), and the per-GPU node tables are merged (nodes above the shard
 * depth span GPUs: their parts are concatenated in rank = Morton order).  RANDOM_GRID / GRID_CENTER / JITTERED
 * results are bit-identical to swgpu_index_batch on one GPU; MIN_DISTANCE as documented at swgpu_set_shard_faces.
 * `devices` may name a GPU more than once (the ranks then share it).  Single batch per run (the device-resident
 * node store of swgpu_set_multi_batch is per GPU and not combined with sharding).
 * attr_host / attr_bytes: optional attribute record per point (4, 8, 12 or 16 bytes) that travels with the point
 * through the exchange (NULL / 0 = none). */
typedef struct swgpu_multi* swgpu_multi_handle;
int swgpu_multi_create(const sw_params* params, const int* devices, uint32_t n_devices, swgpu_multi_handle* out);
void swgpu_multi_destroy(swgpu_multi_handle m);
const char* swgpu_multi_last_error(swgpu_multi_handle m);
/* xyz_host: the whole batch (n x 3 doubles, clamped in place like index_point does) */
int swgpu_multi_index_batch(swgpu_multi_handle m, double* xyz_host, uint64_t n, const void* attr_host,
                            uint32_t attr_bytes);
/* FAST reconstruct on every GPU, then the merge; must be called before the result accessors (ACCURATE too) */
int swgpu_multi_finalize(swgpu_multi_handle m);
int swgpu_multi_result_size(swgpu_multi_handle m, uint64_t* n_nodes, uint64_t* n_point_ids);
/* merged node table + node-major point ids (indices into the batch) */
int swgpu_multi_get_nodes(swgpu_multi_handle m, sw_node* nodes, uint32_t* point_ids);
/* start level (FAST), shard prefix depth, clamped points, points per GPU after the exchange (n_devices entries);
 * any pointer may be NULL */
int swgpu_multi_get_info(swgpu_multi_handle m, int32_t* start_level, uint32_t* shard_levels, uint64_t* n_clamped,
                         uint64_t* shard_points);
/* one GPU's part of the result with the node-major attribute records of its points, gathered on that GPU from
 * the attributes that arrived with the points (swgpu_gather_attribute_device) */
int swgpu_multi_rank_result_size(swgpu_multi_handle m, uint32_t rank, uint64_t* n_nodes, uint64_t* n_point_ids);
int swgpu_multi_get_rank_attributes(swgpu_multi_handle m, uint32_t rank, sw_node* nodes, uint32_t* point_ids,
                                    void* attr_host);
/* MIN_DISTANCE on nodes that span GPUs: 1 (default) = resolve the shard faces (swgpu_set_shard_faces), 0 = off */
int swgpu_multi_set_min_distance_faces(swgpu_multi_handle m, int enable);

/* Algorithmic bytes of the last index_batch + finalize by the accounting model of SURVEY.md section 8(d)
 * (bytes_index/sort/gather/sample; used by bench.py for the roofline line), the bytes this implementation
 * actually reads and writes by its own traffic model (bytes_traffic: e.g. the two-pass compaction reads
 * the keys twice), and the per-stage device times in ms when timing was enabled with swgpu_enable_timing
 * (cudaEvents on the handle's stream). */
typedef struct swgpu_stats {
  uint64_t n_points;
  uint64_t n_output_ids;
  uint64_t n_nodes;
  uint32_t n_levels;     /* sweep levels executed */
  uint32_t n_reconstruct_levels;
  uint64_t sweep_points; /* sum over levels of the list length entering the level */
  uint64_t bytes_index;
  uint64_t bytes_sort;
  uint64_t bytes_gather;
  uint64_t bytes_sample;
  float ms_index;
  float ms_sort;
  float ms_gather;
  float ms_sample;
  float ms_total;
  uint32_t kernel_launches;
  uint32_t min_distance_rounds;
  uint64_t bytes_traffic;
  /* K2 of the last batch (swgpu_set_sort_mode) */
  uint32_t sort_passes;     /* onesweep passes executed (8 = plain LSD; 5 or 6 + the segment finish otherwise) */
  uint32_t sort_first_bit;  /* key bits below this one were ordered by the segment finish kernel (0 = none) */
  uint32_t sort_fallback;   /* runs of equal top bits longer than the finish kernel handles (256) that were not in
                             * order: 2 = each sorted on its own (long_run_sort_kernel), 1 = they held more than 1/8 of
                             * the batch, eight LSD passes were run instead; 0 = none */
  float ms_sort_finish;     /* part of ms_sort spent in the segment finish kernel */
  uint64_t sort_scan_steps; /* neighbour comparisons of the segment finish kernel */
  uint64_t sort_moved;      /* elements the segment finish kernel moved */
} swgpu_stats;
int swgpu_enable_timing(swgpu_handle h, int enable);
int swgpu_get_stats(swgpu_handle h, swgpu_stats* out);

#ifdef __cplusplus
}
#endif
#endif
