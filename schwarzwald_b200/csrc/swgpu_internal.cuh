// swgpu_internal.cuh — host-callable launchers implemented in the kernel translation units.
// All launchers enqueue on `stream` and never synchronise; errors are reported through
// cudaGetLastError() by the caller.
#pragma once

#include "common.cuh"

#include "../../include/sw_types.h"

// ---- index + sort (kernels_index_sort.cu) ---------------------------------------------------
// K1: index_point<21> (clamp + Morton encode) fused with the 8 digit histograms of the sort.
//   xyz        AoS n x 3 doubles (clamped in place when a point lies outside the bounds)
//   keys       n x u64
//   hist       sort_hist_words() u32 (one row of digit counts per radix pass), zeroed by the caller
//   n_clamped  device counter of points that were clamped (zeroed by the caller)
void launch_morton_encode(double* xyz, u64 n, const SwBounds& b, u64* keys, u32* hist, u32* n_clamped,
                          cudaStream_t stream);

// histogram only (keys already exist: tests, multi-GPU path after the shuffle)
void launch_key_histogram(const u64* keys, u64 n, u32* hist, cudaStream_t stream);

// Stable LSD radix sort of (key, id) pairs over the 63 bits of a MortonIndex64 (bit 63 must be clear):
// 7 passes of 9 bits (onesweep: one read + one write per pass, decoupled look-back between tiles).
// `hist` are the digit counts of the keys (sort_hist_words() u32, filled by K1 / launch_key_histogram).
// The unsorted keys must be in keys<sort_input_buffer()>; the sorted result ends in keys0 / vals0.  The
// vals buffers need not be initialised: the first pass generates ids 0..n-1 on the fly.
// `status` must hold sort_status_words(n) u32; `ticket` 8 u32.
size_t sort_hist_words();
int sort_input_buffer();
int sort_passes();
size_t sort_status_words(u64 n);
void launch_radix_sort(u64* keys0, u64* keys1, u32* vals0, u32* vals1, u64 n, u32* hist, u32* status, u32* ticket,
                       cudaStream_t stream);

// The same order with fewer passes over the data (see kernels_index_sort.cu): onesweep passes first_pass .. 7 sort by
// the key bits from 8 * first_pass up (unsorted keys in keys<sort_input_buffer_top(first_pass)>), then the runs of
// equal top bits are ordered in place by their low bits.  first_pass = 0 is launch_radix_sort.
// `stats` (6 u32, 8-byte aligned, zeroed by the caller): [0] bit 31: a run longer than the finish kernel handles was
// not in order, bits 0..30: elements in such long runs; [1] number of long runs, whose first positions the finish
// kernel left in `status` (free after the passes) -> the caller runs launch_long_run_sort(status, stats[1]) or, when
// long runs hold a large part of the points, launch_radix_sort_again (all eight passes over the current arrangement,
// `hist` already scanned); [2..3] u64 scan steps, [4..5] u64 elements moved.  `before_finish`: optional event
// recorded between the passes and the finish kernel.
int sort_input_buffer_top(int first_pass);
void launch_radix_sort_top(u64* keys0, u64* keys1, u32* vals0, u32* vals1, u64 n, int first_pass, u32* hist,
                           u32* status, u32* ticket, u32* stats, cudaStream_t stream, cudaEvent_t before_finish);
void launch_long_run_sort(u64* keys0, u64* keys1, u32* vals0, u32* vals1, u64 n, int first_pass, const u32* long_runs,
                          u32 n_runs, cudaStream_t stream);
// run-length counters of sorted keys (16 u64 + the number of keys sampled, see run_stats_kernel): input of the
// sort-mode choice
void launch_run_stats(const u64* sorted_keys, u64 n, unsigned long long* out17, cudaStream_t stream);
void launch_radix_sort_again(u64* keys0, u64* keys1, u32* vals0, u32* vals1, u64 n, const u32* scanned_hist,
                             u32* status, u32* ticket, cudaStream_t stream);

// positions gathered into Morton order: dst[i] = src[perm[i]] (24-byte records)
void launch_gather_positions(const double* src, const u32* perm, u64 n, double* dst, cudaStream_t stream);
// generic attribute gather for records of `width` bytes (1, 2, 3, 4, 8, 12)
void launch_gather_bytes(const void* src, const u32* perm, u64 n, u32 width, void* dst, cudaStream_t stream);
// out[i] = perm[idx[i]]
void launch_compose_ids(const u32* perm, const u32* idx, u64 n, u32* out, cudaStream_t stream);

// FAST start-level estimate support: bin_start[b] = lower_bound(sorted keys, b << 45) for the
// 8^6 level-5 prefixes (+1 sentinel = n).  TilingAlgorithms.cpp:1473-1535 only needs the sizes of
// the non-empty ranges at levels 0..5, which are sums of these bins.
void launch_level5_bins(const u64* sorted_keys, u64 n, u32* bin_start /* 262145 */, cudaStream_t stream);

// ---- sampling sweep (kernels_sampling.cu) -----------------------------------------------------
#ifndef SWP_THREADS
#define SWP_THREADS 256
#endif
#define SWP_ITEMS 8
#define SW_SWEEP_TILE (SWP_THREADS * SWP_ITEMS) /* elements per tile of the rle / select / compact kernels */

struct SwLevelArgs
{
  // input list (Morton ordered)
  const u64* in_key; // bit 63 clear
  const u32* in_idx; // index into the sorted order; nullptr = identity
  u64 count;
  // node definition: nodes are runs of equal (key >> node_shift)
  int node_shift;
  // sampling cell definition for the grid strategies: runs of equal (key >> cell_shift)
  int cell_shift;
  // behaviour
  int sampling;          // sw_sampling
  int force_all;         // terminal level: every point is taken
  int allow_take_all;    // TakeAllWhenCountBelowMaxPoints (1) or AlwaysAdhereToMinSpacing (0)
  u64 max_points_per_node;
  // node table produced by the run-length encoder
  const u32* node_start; // n_nodes + 1 entries (sentinel = count)
  const u32* tile_rank0; // per tile: number of node heads before the tile
  const u32* node_gcount; // optional: global (all-shard) point count per node rank, nullptr = local run length
  // optional per-element selection flags (argmin / min-distance strategies)
  const unsigned char* sel;
  // scratch of the two-pass compaction
  u64* status;      // single-pass compaction: sweep_tiles look-back descriptors (nullptr = two-pass kernels)
  u32* ticket;      // single-pass compaction: tile ticket
  u32* selbits;     // one bit per element: sweep_tiles * SW_SWEEP_TILE / 32 words
  u32* tile_sel;    // per tile: selected points (count pass), then their exclusive scan
  u32* child_count; // 8 * n_nodes counters of the points that stay, per child node; nullptr = not needed
  // outputs
  u64* out_key;
  u32* out_idx;
  u64 out_offset; // where this level's chunk starts in out_key / out_idx
  u64* rem_key;   // nullptr = remainder is dropped (reconstruct)
  u32* rem_idx;
  // node table output (indexed by node rank + node_base)
  u64* node_index;
  u64* node_first;
  u64 node_base;
  int levels; // number of levels of the nodes at this level (node level + 1)
};

size_t sweep_tiles(u64 count);
// run-length encode nodes: node_start[], tile_rank0[], *n_nodes.  `status` >= sweep_tiles u64.
void launch_node_rle(const u64* keys, u64 count, int node_shift, u32* node_start, u32* tile_rank0, u32* n_nodes,
                     u64* status, u32* ticket, cudaStream_t stream);
// node boundaries that do not need a pass over the points
void launch_root_node(u32* node_start, u64 count, cudaStream_t stream);
// FAST start nodes from the level-5 bin boundaries (launch_level5_bins) of the sorted keys
void launch_start_nodes(const u32* bin_start, int start_levels, u64 count, u32* node_start, u32* n_nodes,
                        cudaStream_t stream);
// reconstruct: parent nodes = runs of children (index >> 3) in the previous output chunk
void launch_parent_nodes(const u64* child_index, const u64* child_first, u32 n_children, u64 chunk_offset,
                         u64 chunk_count, u32* node_start, u32* n_nodes, cudaStream_t stream);
void launch_tile_rank0(const u32* node_start, u32 n_nodes, u64 count, u32* tile_rank0, cudaStream_t stream);
// Stable two-way compaction of one level (count pass, scan, scatter pass).  *n_selected receives the
// number of selected points.  With a.child_count != nullptr the node boundaries of the next level
// are written to next_node_start (n_child_slots = 8 * nodes of this level) and their number to
// *n_nodes_next.
void launch_level_compact(const SwLevelArgs& a, u64* n_selected, u32 n_child_slots, u32* next_node_start,
                          u32* n_nodes_next, cudaStream_t stream);

// per-node quantities of the GRID_CENTER / JITTERED selection (argmin_nodes_kernel)
struct SwArgminNode
{
  double mn[3], mx[3]; // node bounds by the get_octant_bounds recurrence
  double grid_cell_size, permutation_cell_size; // JITTERED, Sampling.h:655-660
  int shift;  // key shift of the selection cells
  int levels; // JITTERED: log2(cells per axis)
  u32 cells;
  u32 pad;
};

struct SwArgminArgs
{
  const u64* in_key;
  const u32* in_idx; // nullptr = identity
  u64 count;
  const double* pos_sorted; // AoS positions in Morton order
  int sampling;             // SW_GRID_CENTER or SW_JITTERED
  int node_shift;
  int node_level; // reference node level (root = -1)
  int cell_shift; // GRID_CENTER: shift of the candidate level; JITTERED: computed per node
  int cand_level; // GRID_CENTER candidate level c (cell bounds at depth c + 1)
  double spacing_at_node;
  SwBounds bounds;
  unsigned char* sel; // zeroed by the caller; 1 = winner of its cell
  u32* error_flag;    // JITTERED: set to SW_ERR_JITTER_* by the kernel
  // take-all nodes are skipped (they are never sampled: Sampling.h:328-335, 612-619)
  const u32* node_start;
  const u32* tile_rank0;
  const u32* node_gcount; // see SwLevelArgs
  int allow_take_all;
  u64 max_points_per_node;
  u32 n_nodes;
  SwArgminNode* nodes; // n_nodes entries, filled by launch_select_argmin
};
// status >= 5 * sweep_tiles(count) u64
void launch_select_argmin(const SwArgminArgs& a, u64* status, u32* ticket, cudaStream_t stream);

// grow-only device buffer owned by the caller (freed by the tiler handle)
struct SwGrowBuf
{
  void* p;
  size_t cap;
};

struct SwMinDistArgs
{
  const u64* in_key;
  const u32* in_idx; // nullptr = identity
  u64 count;
  const double* pos_sorted;
  int node_shift;
  int node_levels;  // levels of the nodes of this sweep level (node level + 1)
  int cell_levels;  // levels of the neighbour-search cells (cell side >= spacing), <= 21
  double threshold; // (double)(float)(spacing_f * spacing_f), SparseGrid.cpp:11-14
  const u32* node_start;
  const u32* tile_rank0;
  const u32* node_gcount; // see SwLevelArgs
  int allow_take_all;
  u64 max_points_per_node;
  u32 nth_point; // 1 = every point is analysed; n = only every n-th point of a node (MIN_DISTANCE_FAST)
};

struct SwMinDistScratch
{
  SwGrowBuf cell_start, cell_tile_rank0, state, lpos, desc, acc_xyz, hkeys, hvals, hmask, nbr, deps, queue, cell_active,
    counters;
  u64* status; // look-back descriptors (>= sweep_tiles u64)
  u32* ticket;
  u32* h_pinned; // pinned host scratch, >= 8 u32
};

// Greedy minimum-distance selection for every sampled node of one level.  On return
// scratch.state.p holds one byte per point: 1 = accepted, 2 = rejected / not sampled.
cudaError_t run_min_distance(const SwMinDistArgs& a, SwMinDistScratch& sc, cudaStream_t stream, u32* rounds,
                             u32* launches, u64* bytes);
void free_min_distance_scratch(SwMinDistScratch& sc);

// ---- multi-GPU support (kernels_shard.cu) ---------------------------------------------------------
#define SW_MAX_RANKS 16
#define SW_PREFIX_BINS 262144 /* 8^6 level-5 prefixes = key >> 45 */

// bins[key >> 45] += 1 for unsorted keys (bins are accumulated, not zeroed)
void launch_prefix_histogram(const u64* keys, u64 n, u32* bins, cudaStream_t stream);
// the same for the leading `levels` <= 4 octree levels (8^levels bins), privatised in shared memory
void launch_prefix_histogram_coarse(const u64* keys, u64 n, int levels, u32* bins, cudaStream_t stream);
// counts[b] = bin_start[b + 1] - bin_start[b]
void launch_bin_counts(const u32* bin_start, u32 n_bins, u32* counts, cudaStream_t stream);
// Stable multi-way partition of (xyz, id_base + i) by destination rank; rank r owns the level-5
// prefixes [first_prefix[r], first_prefix[r+1]).  tile_counts: partition_tiles(n) * SW_MAX_RANKS
// u32 scratch; send_counts: SW_MAX_RANKS u64 (device).
size_t partition_tiles(u64 n);
// attr / attr_words / out_attr: optional attribute record of attr_words (1..4) 32-bit words per point that is
// partitioned along with the positions (nullptr = none)
void launch_partition_by_splitters(const u64* keys, const double* xyz, u64 n, const u32* first_prefix, u32 n_ranks,
                                   u32 id_base, u32* tile_counts, u64* send_counts, double* out_xyz, u32* out_id,
                                   const u32* attr, u32 attr_words, u32* out_attr, cudaStream_t stream);
// Same partition, written straight into the destinations' receive buffers: peer_xyz[r] / peer_ids[r]
// are device pointers to rank r's buffers (peer-mapped over NVLink, or local), dst_offsets[r] is the
// first point of this source's block there.  send_counts (device, SW_MAX_RANKS u64) still receives
// the per-destination totals.
void launch_partition_to_peers(const u64* keys, const double* xyz, u64 n, const u32* first_prefix, u32 n_ranks,
                               u32 id_base, u32* tile_counts, u64* send_counts, double* const* peer_xyz,
                               u32* const* peer_ids, const u64* dst_offsets, const u32* attr, u32 attr_words,
                               u32* const* peer_attr, cudaStream_t stream);
// node counts of one sweep level <-> dense per-prefix counters (8^levels entries)
void launch_node_counts_to_dense(const u64* keys, const u32* node_start, u32 n_nodes, int node_shift, u32* dense,
                                 cudaStream_t stream);
void launch_node_counts_from_dense(const u64* keys, const u32* node_start, u32 n_nodes, int node_shift,
                                   const u32* dense, u32* gcount, cudaStream_t stream);
// out[i] = map[perm[idx[i]]]
void launch_compose_ids_mapped(const u32* perm, const u32* idx, const u32* map, u64 n, u32* out, cudaStream_t stream);

// MIN_DISTANCE across shard faces (see kernels_shard.cu): accepted points that have another shard within the
// spacing, exchanged between the ranks; a point loses against a conflicting accepted point of a lower rank
struct SwFaceRecord
{
  u64 key;
  double x, y, z;
};
struct SwFaceRanks
{
  u64 first[SW_MAX_RANKS + 1]; // records of rank r = [first[r], first[r + 1]) of the gathered array
};
void launch_face_flag(const u32* in_idx, const double* pos, const unsigned char* state, u64 count, const SwBounds& b,
                      double reach, const u32* first_prefix, u32 n_ranks, u32 my_rank, u32* flags, cudaStream_t stream);
void launch_face_collect(const u64* in_key, const u32* in_idx, const double* pos, u64 count, const u32* flags,
                         const u64* offs, SwFaceRecord* rec, u32* rec_src, cudaStream_t stream);
void launch_face_resolve(const SwFaceRecord* mine, const u32* mine_src, u32 n_mine, const SwFaceRecord* all,
                         const SwFaceRanks& fr, u32 my_rank, int cell_levels, int node_levels, double threshold,
                         unsigned char* state, cudaStream_t stream);

// ---- LAS input transform (kernels_index_sort.cu) and writer payloads (kernels_payload.cu) ------------
// device-side copy of sw_las_transform (include/sw_types.h)
struct SwLasTransform
{
  double scale[3];
  double offset[3];
  double hmin[3];
  double hmax[3];
  double center[3];
  int shift;
};
// K1-LAS: position_from_las_point (io/LASFile.cpp:79-94) + shift/float rounding
// (process/TilerProcess.cpp:552-559) + index_point, fused; same outputs as launch_morton_encode plus
// the positions themselves (xyz_out, n x 3 doubles)
void launch_las_encode(const int* las, u64 n, const SwLasTransform& t, const SwBounds& b, double* xyz_out, u64* keys,
                       u32* hist, u32* n_clamped, cudaStream_t stream);

// PNTS payload: out[j] = (float) position of the j-th node-major point (io/PNTSWriter.cpp:326-342).
// perm == nullptr: `xyz` is already in sorted order (indexed by out_idx directly).
void launch_payload_pnts(const double* xyz, const u32* perm, const u32* out_idx, u64 n_out, float* out,
                         cudaStream_t stream);
// LAS payload: per node header offset = node bounds min, one scale per node; X = I32_QUANTIZE((p - offset) /
// scale) (io/LASPersistence.h:119-131,160-163 + LASzip's laszip_set_coordinates).  node_first: first
// output offset per node (bit 63 may carry a flag); node_hdr: 4 doubles per node (offset xyz, scale).
void launch_payload_las(const double* xyz, const u32* perm, const u32* out_idx, u64 n_out, const u64* node_first,
                        u32 n_nodes, const double* node_hdr, int* out, cudaStream_t stream);

// ---- multi-batch node store (kernels_store.cu) ------------------------------------------
size_t scan_scratch_words(u64 n);
// out[0..n] = exclusive scan of in[0..n) as u64 (out[n] = total); scratch: scan_scratch_words(n) u64
void launch_exclusive_scan_u32(const u32* in, u64 n, u64* out, u64* scratch, cudaStream_t stream);
// gid[i] = base + order[i]
void launch_make_gids(const u32* order, u64 n, u32 base, u32* gid, cudaStream_t stream);
// visited nodes (runs of the incoming list) -> slot in the level's sorted node table (0xFFFFFFFF: new node) and
// the number of points stored there
void launch_store_lookup(const u64* in_key, const u32* node_start, u32 n_nodes, int node_shift, const u64* st_index,
                         u32 st_n, const u64* st_first, u32* slot, u32* cnt, cudaStream_t stream);
// the stored points of the visited nodes as a (key, gid) list, keys re-derived relative to the node bounds
// (read_pnts_from_disk, TilingAlgorithms.cpp:50-109); boff = exclusive scan of the stored counts (n_nodes + 1)
void launch_store_fetch(u64 m, const u64* boff, u32 n_nodes, const u32* slot, const u64* st_first, const u32* st_ids,
                        const u64* in_key, const u32* node_start, int levels, const double* xyz, const SwBounds& root,
                        u64* bkey, u32* bidx, cudaStream_t stream);
// keys of stored ids relative to the root bounds (reconstruct input)
void launch_store_root_keys(const u32* ids, u64 n, const double* xyz, const SwBounds& root, u64* keys,
                            cudaStream_t stream);
void launch_store_merged_nodes(const u32* node_start_a, const u64* boff, u32 n_nodes, u32* node_start_c, u32* gcount,
                               cudaStream_t stream);
// C = merge(A, B) by key, A first on ties (merge_node_data_sorted, Node.cpp:3-20)
void launch_merge_lists(const u64* ak, const u32* ai, u64 na, const u64* bk, const u32* bi, u64 nb, u64* ck, u32* ci,
                        cudaStream_t stream);
// per node: A's run, then B's run (merge_node_data_unsorted, Node.cpp:22-34)
void launch_concat_lists(const u64* ak, const u32* ai, u64 na, const u64* bk, const u32* bi, u64 nb,
                         const u32* node_start_a, const u64* boff, u32 n_nodes, u64* ck, u32* ci, cudaStream_t stream);
// store update (see kernels_store.cu)
void launch_store_match(const u64* vidx, u32 nv, const u64* oidx, u32 no, u32* lo, u32* found, cudaStream_t stream);
void launch_store_rows(const u64* vidx, const u64* vfirst, u32 nv, u64 chunk_end, u32 chunk_flags, const u32* lo,
                       const u64* cumf, const u64* oidx, const u64* ofirst, const u32* oflags, u32 no, u64* nidx,
                       u32* ncnt, u32* nflags, u64* nsrc, cudaStream_t stream);
void launch_store_copy(u64 total, const u64* nfirst, u32 nn, const u64* nsrc, const u32* old_ids, const u32* chunk_ids,
                       u32* new_ids, cudaStream_t stream);
