// tiler.cu — host-side orchestration of the device pipeline and the C ABI of include/swgpu.h.
//
// One batch (the parity regime of SURVEY.md: internal_cache_size >= N):
//   index (K1) -> sort (K2) -> [gather positions (K4)] -> level-synchronous sampling sweep
//   ACCURATE (TilingAlgorithmV1, tiling/TilingAlgorithms.cpp:577-626): sweep starts at the root
//            (node level -1, one node holding every point)
//   FAST     (TilingAlgorithmV3, :1250-1360): start level S from the sizes of the sorted key
//            ranges (:1473-1535), sweep starts at node level S-1 (nodes with S levels);
//            swgpu_finalize() re-samples levels S-1..0 from their children (:1661-1784)
//
// Per-level decisions that the reference takes per node but that only depend on the node LEVEL are
// tabulated on the host with the reference's exact float/double narrowing
// (tiling/Node.cpp:37-57, tiling/Sampling.cpp:29-62, tiling/Sampling.h:210-229).
#include "swgpu_internal.cuh"

#include "../../include/swgpu.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

// SWGPU_TRACE_ALLOC=1: every (re)allocation of a device buffer is reported on stderr with its duration
static bool
trace_alloc()
{
  static int on = -1;
  if (on < 0) {
    const char* e = std::getenv("SWGPU_TRACE_ALLOC");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

struct DevBuf
{
  void* p = nullptr;
  size_t cap = 0;

  // Grows the buffer to at least `bytes`; optionally preserves `keep_bytes` of content.
  cudaError_t ensure(size_t bytes, cudaStream_t stream = nullptr, size_t keep_bytes = 0)
  {
    if (bytes <= cap)
      return cudaSuccess;
    const auto t0 = std::chrono::steady_clock::now();
    struct Report
    {
      std::chrono::steady_clock::time_point t0;
      size_t from, to, keep;
      ~Report()
      {
        if (trace_alloc())
          std::fprintf(stderr, "[swgpu alloc] %zu -> %zu bytes (keep %zu): %.3f ms\n", from, to, keep,
                       std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      }
    } report{ t0, cap, bytes, keep_bytes };
    size_t want = bytes;
    if (keep_bytes) // amortise growth of append-only buffers
      want = std::max(bytes, cap + cap / 2);
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess)
      return e;
    if (p && keep_bytes) {
      e = cudaMemcpyAsync(np, p, keep_bytes, cudaMemcpyDeviceToDevice, stream);
      if (e == cudaSuccess)
        e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess) {
        cudaFree(np);
        return e;
      }
    }
    if (p)
      cudaFree(p);
    p = np;
    cap = want;
    return cudaSuccess;
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template<typename T>
  T* as() const
  {
    return static_cast<T*>(p);
  }
};

struct Chunk
{
  int levels;     // number of levels of the nodes in this chunk (node level + 1)
  u64 out_offset; // first id of the chunk in the output arrays
  u64 count;      // ids in the chunk
  u64 node_base;  // first row in the node table
  u32 n_nodes;
  u32 flags;
};

struct HostScalars
{
  u32 n_nodes;
  u32 n_clamped;
  u32 error_flag;
  u32 n_nodes_next;
  u64 n_selected;
  u32 scratch[16]; // pinned scratch for the min-distance driver (starts at u32 index 8... see below)
  u64 store_vals[4]; // read-backs of the multi-batch node store
  u32 sort_stats[6]; // read-back of d_sort_stats(): flag + long elements, long runs, u64 scan steps, u64 moved
  u64 run_stats[17]; // read-back of run_stats_kernel
};

// One octree level of the multi-batch node store: node table sorted by node index and the
// stored global point ids, node by node, in stored order.
struct LevelStore
{
  DevBuf index; // u64 n_nodes
  DevBuf first; // u64 n_nodes + 1 (exclusive scan of the counts)
  DevBuf flags; // u32 n_nodes
  DevBuf ids;   // u32 n_ids
  // the rebuild of store_update writes into the level's own spare set and swaps: the sizes of a level only
  // grow, so after a few batches no allocation happens any more (a set shared by all levels would be
  // re-allocated on nearly every swap)
  DevBuf spare_index, spare_first, spare_flags, spare_ids;
  u64 n_nodes = 0;
  u64 n_ids = 0;
};

// grow with head room (append-only tables): at least `bytes`, twice the old capacity when it has to grow
cudaError_t
ensure_amortised(DevBuf& b, size_t bytes)
{
  if (bytes <= b.cap)
    return cudaSuccess;
  return b.ensure(std::max(bytes, 2 * b.cap));
}

enum LevelKind
{
  KIND_INTERNAL = 0,
  KIND_TERMINAL = 1,
  KIND_REROOT = 2
};

} // namespace

struct swgpu_tiler
{
  sw_params prm{};
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  SwBounds bounds{};

  // current batch
  double* d_xyz = nullptr; // owned (xyz_own) or borrowed
  u64 n = 0;
  bool batch_done = false;
  bool finalized = false;
  int32_t start_level = -1;
  u64 n_clamped = 0;

  DevBuf xyz_own;
  const int* d_las = nullptr; // current batch arrives as LAS record coordinates (swgpu_index_batch_las*)
  SwLasTransform las_t{};
  DevBuf las_own, payload_tmp, node_hdr, argmin_nodes;
  DevBuf keys[2], vals[2];
  DevBuf wkey2, widx2;
  DevBuf hist, sort_status, scalars;
  DevBuf pos_sorted;
  DevBuf out_key, out_idx;
  u64 out_count = 0;
  DevBuf node_start, node_start_next, tile_rank0, sel, scan_status;
  DevBuf selbits, tile_sel, child_count;
  DevBuf node_index, node_first;
  u64 node_count = 0;
  DevBuf bins;
  DevBuf ids_tmp;
  SwMinDistScratch md{}; // min-distance scratch

  HostScalars* h_scalars = nullptr; // pinned
  std::vector<Chunk> chunks;

  // multi-GPU: this handle tiles one shard (whole Morton-prefix subtrees)
  u32 shard_levels = 0;             // nodes with fewer levels may span GPUs
  int32_t start_level_override = -1; // FAST: the global start level
  swgpu_allreduce_u32_fn allreduce = nullptr;
  void* allreduce_ctx = nullptr;
  const u32* global_ids = nullptr; // device: received point -> global point id
  DevBuf dense_counts, node_gcount;
  DevBuf part_tile_counts, part_send_counts;
  // attribute record that travels with the points through the next partition call (swgpu_set_partition_attributes)
  const void* part_attr_src = nullptr;
  u32 part_attr_words = 0;
  void* part_attr_dst = nullptr;          // swgpu_partition_device: send buffer
  void* const* part_attr_peers = nullptr; // swgpu_partition_to_peers_device: receive buffers of all ranks
  // MIN_DISTANCE across shard faces (swgpu_set_shard_faces)
  swgpu_allgatherv_fn face_fn = nullptr;
  void* face_ctx = nullptr;
  u32 face_first_prefix[SW_MAX_RANKS + 1] = {};
  u32 face_n_ranks = 0, face_rank = 0;
  DevBuf face_flags, face_offs, face_scan, face_rec, face_src;
  u64 face_points_sent = 0, face_points_rejected_upper_bound = 0;

  int deep_policy = 0; // swgpu_set_deep_node_policy
  // multi-batch mode (swgpu_set_multi_batch): the node store that persists between batches
  bool multi_batch = false;
  DevBuf store_xyz;       // positions of every batch, clamped; global point id = row
  u64 store_points = 0;
  int32_t store_start_level = -1; // FAST: fixed by the first batch
  LevelStore store[22];
  DevBuf list_key[4], list_idx[4]; // (key, gid) lists of the sweep: input, fetched, merged, remainder
  DevBuf st_slot, st_cnt, st_boff, st_scan, st_gcount, st_lo, st_found, st_cumf;
  DevBuf st_nidx, st_ncnt, st_nflags, st_nsrc, st_nfirst, st_nids;
  // RANDOM_GRID never reads positions: its sweep lists carry the ORIGINAL point index (the sort's permutation)
  // instead of the sorted position, so the output ids need no composition with the permutation afterwards
  bool out_idx_original = false;
  const double* sample_pos = nullptr; // positions the selection kernels read, indexed by the lists' ids
  const u32* gcount_override = nullptr;

  // stats
  swgpu_stats stats{};
  bool timing = false;
  cudaEvent_t ev[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }; // [6]: before the sort's finish kernel

  // K2 (swgpu_set_sort_mode): passes over the key bits from 8 * sort_first_pass up, then the segment finish
  bool two_pass_compaction = true; // SWGPU_COMPACT=1pass selects level_compact_fused_kernel (A/B measurements)
  int sort_mode = -1;      // -1 = automatic
  int sort_first_pass = 0; // of the current batch
  int sort_next = 0;       // automatic mode: first pass of the next batch (0 until a probe found short runs)
  int sort_probe_wait = 0; // automatic mode, eight passes: batches until the sorted keys are probed again

  // device scalar slots inside `scalars`
  u32* d_n_nodes() { return scalars.as<u32>() + 0; }
  u32* d_n_clamped() { return scalars.as<u32>() + 1; }
  u32* d_error() { return scalars.as<u32>() + 2; }
  u32* d_n_nodes_next() { return scalars.as<u32>() + 3; }
  u64* d_n_selected() { return reinterpret_cast<u64*>(scalars.as<u32>() + 4); }
  u32* d_tickets() { return scalars.as<u32>() + 8; } // 16 u32
  u32* d_md_ncells() { return scalars.as<u32>() + 24; }
  u32* d_sort_stats() { return scalars.as<u32>() + 26; } // 6 u32, 8-byte aligned
};

namespace {

#define CK(expr)                                                                                                       \
  do {                                                                                                                 \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess)                                                                                             \
      return fail_cuda(h, _e, #expr);                                                                                  \
  } while (0)

int
fail(swgpu_tiler* h, int code, const std::string& msg)
{
  h->err = msg;
  return code;
}

int
fail_cuda(swgpu_tiler* h, cudaError_t e, const char* what)
{
  h->err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  cudaGetLastError();
  return e == cudaErrorMemoryAllocation ? SW_ERR_OUT_OF_MEMORY : SW_ERR_CUDA;
}

bool
needs_positions(int sampling)
{
  return sampling != SW_RANDOM_GRID;
}

double
root_extent_x(const swgpu_tiler* h)
{
  return h->prm.bounds_max[0] - h->prm.bounds_min[0];
}

// candidate level used inside sample_points: Sampling.h:210-229 (double spacing)
int
cand_level_sampler(const swgpu_tiler* h, int node_level)
{
  const double spacing_at_this_node = h->prm.spacing_at_root / std::pow(2, node_level + 1);
  const float ratio = static_cast<float>(root_extent_x(h) / spacing_at_this_node);
  return std::max(-1, static_cast<int>(std::floor(std::log2f(ratio))) - 1);
}

uint32_t
prev_pow2_u32(uint32_t x)
{
  x |= x >> 1;
  x |= x >> 2;
  x |= x >> 4;
  x |= x >> 8;
  x |= x >> 16;
  return x - (x >> 1);
}

// required_morton_index_depth with the real root (Sampling.cpp:29-62, Node.cpp:37-57)
int
required_depth(const swgpu_tiler* h, int node_level)
{
  switch (h->prm.sampling) {
    case SW_RANDOM_GRID:
    case SW_GRID_CENTER: {
      const double spacing_at_target = h->prm.spacing_at_root / std::pow(2, node_level + 1);
      const float target_spacing = static_cast<float>(spacing_at_target); // narrowed PARAMETER
      const float ratio = static_cast<float>(root_extent_x(h) / target_spacing);
      return std::max(-1, static_cast<int>(std::floor(std::log2f(ratio))) - 1);
    }
    case SW_MIN_DISTANCE:
    case SW_MIN_DISTANCE_FAST:
      return node_level;
    case SW_JITTERED: {
      const double spacing_at_this_node = h->prm.spacing_at_root / std::pow(2, node_level + 1);
      const double perfect = (root_extent_x(h) / std::pow(2, node_level + 1)) / spacing_at_this_node;
      const uint32_t cells = prev_pow2_u32(static_cast<uint32_t>(perfect));
      const uint32_t levels = cells ? static_cast<uint32_t>(std::log2(cells)) : 0u;
      return static_cast<int32_t>(static_cast<uint32_t>(node_level + levels));
    }
  }
  return node_level;
}

// tile_node branching, TilingAlgorithms.cpp:406-491
LevelKind
level_kind(const swgpu_tiler* h, int node_level)
{
  const int sample_level = required_depth(h, node_level);
  const bool requires_deeper = sample_level > node_level;
  const int max_level = static_cast<int>(std::min<uint32_t>(20u, h->prm.max_depth));
  if (!requires_deeper)
    return sample_level >= max_level ? KIND_TERMINAL : KIND_INTERNAL;
  if (node_level >= max_level)
    return KIND_TERMINAL;
  if (sample_level >= 21)
    return KIND_REROOT;
  return KIND_INTERNAL;
}

int
sync_scalars(swgpu_tiler* h)
{
  CK(cudaMemcpyAsync(h->h_scalars, h->scalars.p, 24, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SW_OK;
}

// ---- K2 ------------------------------------------------------------------------------------------------
// Which onesweep pass the sort of this batch starts with (kernels_index_sort.cu: passes over the top digits, then
// the segment finish).  0 = all eight passes.  Automatic mode never guesses: the first batch of a handle is sorted
// by the eight passes and its sorted keys are probed (run_stats_kernel, a read of every 8th chunk of the keys) for the lengths of
// the runs of equal top 40 / 48 bits; later batches start at pass 3 (top 40 bits = 13 octree levels) when those
// runs are short (a terrain model: 1.6 points per run), at pass 2 (48 bits) when only those are, and stay with the
// eight passes for clustered clouds (urban LiDAR: a fifth of the points in runs of more than eight at 6 cm), where
// ordering the runs costs more than the passes it saves.  A batch that turns out denser than the probed one sends
// the handle back to the eight passes and a new probe.  The order produced is the same in every mode.
int
choose_sort_first_pass(const swgpu_tiler* h)
{
  if (const char* e = std::getenv("SWGPU_SORT_FIRST_PASS")) { // experiments only
    const int v = std::atoi(e);
    if (v >= 0 && v <= 3)
      return v;
  }
  if (h->sort_mode >= 0)
    return h->sort_mode;
  return h->sort_next > 0 ? h->sort_next : 0;
}

// bound on the comparisons per point the finish kernel needs for runs of equal key bits >= 24 (g = 0) or 16 (g = 1)
// and the share of points with 128 or more predecessors in their run, from the run_stats_kernel counters
void
run_stats_summary(const u64* c, int g, u64 n, double* steps_per_point, double* long_share)
{
  double sum = 0;
  for (int q = 0; q < 8; ++q)
    sum += (double)(1u << q) * (double)c[g * 8 + q];
  *steps_per_point = 2.0 * sum / (double)n;
  *long_share = (double)c[g * 8 + 7] / (double)n;
}

#define SORT_STEPS_MAX_40 3.0 /* finish kernel at 1.15 steps per point: 0.66 ms per 100 M points, three passes: 2.0 ms */
#define SORT_STEPS_MAX_48 1.0
#define SORT_LONG_SHARE_MAX 1e-3

// sorted pairs end in keys0 / vals0; the unsorted keys are in keys<sort_input_buffer_top(fp)>.  `adapt`: a batch of
// the tiler (automatic mode learns from it), not the stand-alone primitive.
int
sort_pairs(swgpu_tiler* h, u64* keys0, u64* keys1, u32* vals0, u32* vals1, u64 n, int fp, bool adapt)
{
  cudaStream_t s = h->stream;
  launch_radix_sort_top(keys0, keys1, vals0, vals1, n, fp, h->hist.as<u32>(), h->sort_status.as<u32>(),
                        h->d_tickets() + 8, h->d_sort_stats(), s, h->timing ? h->ev[6] : nullptr);
  u32 passes = (u32)(sort_passes() - fp);
  h->stats.kernel_launches += 1 + passes + (fp ? 1 : 0);
  h->stats.sort_first_bit = 8u * (u32)fp;
  h->stats.sort_fallback = 0;
  h->stats.sort_scan_steps = 0;
  h->stats.sort_moved = 0;
  u64 moved = 0;
  adapt = adapt && h->sort_mode < 0 && n > 0;
  if (fp && n) {
    CK(cudaMemcpyAsync(h->h_scalars->sort_stats, h->d_sort_stats(), 24, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const u32* st = h->h_scalars->sort_stats;
    std::memcpy(&h->stats.sort_scan_steps, st + 2, 8);
    std::memcpy(&moved, st + 4, 8);
    h->stats.sort_moved = moved;
    const bool unsorted_long = (st[0] >> 31) != 0;
    const u64 long_elements = st[0] & 0x7fffffffu;
    if (unsorted_long && long_elements * 8 > n) {
      // long runs of equal top bits hold much of the batch: all eight passes over the current arrangement
      launch_radix_sort_again(keys0, keys1, vals0, vals1, n, h->hist.as<u32>(), h->sort_status.as<u32>(),
                              h->d_tickets() + 8, s);
      h->stats.sort_fallback = 1;
      h->stats.kernel_launches += (u32)sort_passes();
      passes += (u32)sort_passes();
    } else if (unsorted_long) { // a few long runs (dense spots): each is sorted on its own
      launch_long_run_sort(keys0, keys1, vals0, vals1, n, fp, h->sort_status.as<u32>(), st[1], s);
      h->stats.kernel_launches += 1;
      h->stats.sort_fallback = 2;
    }
    if (adapt) { // denser than the batch that was probed: back to the eight passes, probe again
      const double steps_per_point = (double)h->stats.sort_scan_steps / (double)n;
      const double limit = fp >= 3 ? SORT_STEPS_MAX_40 : SORT_STEPS_MAX_48;
      if (h->stats.sort_fallback == 1 || steps_per_point > 1.5 * limit ||
          (double)long_elements > 2 * SORT_LONG_SHARE_MAX * (double)n) {
        h->sort_next = 0;
        h->sort_probe_wait = 0;
      }
    }
  } else if (adapt) {
    if (h->sort_probe_wait > 0) {
      --h->sort_probe_wait;
    } else {
      // the histogram rows are free after the passes: 17 u64 counters
      unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(h->hist.p);
      launch_run_stats(keys0, n, d_counts, s);
      h->stats.kernel_launches += 1;
      CK(cudaMemcpyAsync(h->h_scalars->run_stats, d_counts, 17 * 8, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      double steps = 0, long_share = 0;
      const u64 sampled = std::max<u64>(h->h_scalars->run_stats[16], 1);
      run_stats_summary(h->h_scalars->run_stats, 0, sampled, &steps, &long_share);
      if (steps <= SORT_STEPS_MAX_40 && long_share <= SORT_LONG_SHARE_MAX) {
        h->sort_next = 3;
      } else {
        run_stats_summary(h->h_scalars->run_stats, 1, sampled, &steps, &long_share);
        h->sort_next = (steps <= SORT_STEPS_MAX_48 && long_share <= SORT_LONG_SHARE_MAX) ? 2 : 0;
      }
      if (h->sort_next == 0)
        h->sort_probe_wait = 15; // clustered cloud: look again after 15 more batches
    }
  }
  h->stats.sort_passes = passes;
  // SURVEY 8(d) K2, the accounting model of the roofline: histogram read + 8 passes x 2 x 12 B per point, whatever
  // this implementation moves (sort_passes and bytes_traffic say what it did move)
  h->stats.bytes_sort = (8 + (u64)sort_passes() * 24) * n;
  // traffic model: histograms come from K1, the first pass reads no ids, the finish writes what it moves
  h->stats.bytes_traffic += ((u64)passes * 24 - 4 + (fp ? 12 : 0)) * n + 12 * moved;
  return SW_OK;
}

int
sort_batch(swgpu_tiler* h, u64 n)
{
  return sort_pairs(h, h->keys[0].as<u64>(), h->keys[1].as<u64>(), h->vals[0].as<u32>(), h->vals[1].as<u32>(), n,
                    h->sort_first_pass, true);
}

int
ensure_batch_buffers(swgpu_tiler* h, u64 n)
{
  const size_t nn = (size_t)std::max<u64>(n, 1);
  for (int i = 0; i < 2; ++i) {
    CK(h->keys[i].ensure(nn * 8));
    CK(h->vals[i].ensure(nn * 4));
  }
  CK(h->wkey2.ensure(nn * 8));
  CK(h->widx2.ensure(nn * 4));
  CK(h->hist.ensure(sort_hist_words() * 4));
  CK(h->sort_status.ensure(sort_status_words(n) * 4));
  CK(h->node_start.ensure((nn + 1) * 4));
  CK(h->node_start_next.ensure((nn + 1) * 4));
  CK(h->selbits.ensure(sweep_tiles(n) * (SW_SWEEP_TILE / 32) * 4));
  CK(h->tile_sel.ensure(sweep_tiles(n) * 4));
  CK(h->tile_rank0.ensure(sweep_tiles(n) * 4));
  CK(h->scan_status.ensure(sweep_tiles(n) * 8 * 5));
  CK(h->bins.ensure(262145 * 4));
  if (needs_positions(h->prm.sampling)) {
    CK(h->pos_sorted.ensure(nn * 24));
    CK(h->sel.ensure(nn));
  }
  CK(h->out_key.ensure(nn * 8));
  CK(h->out_idx.ensure(nn * 4));
  return SW_OK;
}

void
record(swgpu_tiler* h, int i)
{
  if (h->timing)
    cudaEventRecord(h->ev[i], h->stream);
}

bool
spans_shards(const swgpu_tiler* h, int levels)
{
  return h->allreduce && (u32)levels < h->shard_levels;
}

// Sums the per-node point counts of one sweep level over all shards.  Every rank calls this once
// per level below the shard prefix depth, with n_nodes == 0 when it has nothing left there.
int
exchange_node_counts(swgpu_tiler* h, const u64* in_key, u32 n_nodes, int levels)
{
  const size_t dense_n = (size_t)1 << (3 * levels);
  cudaStream_t s = h->stream;
  CK(h->dense_counts.ensure(dense_n * 4));
  CK(h->node_gcount.ensure((size_t)std::max<u32>(n_nodes, 1) * 4));
  CK(cudaMemsetAsync(h->dense_counts.p, 0, dense_n * 4, s));
  launch_node_counts_to_dense(in_key, h->node_start.as<u32>(), n_nodes, shift_for_levels(levels),
                              h->dense_counts.as<u32>(), s);
  CK(cudaGetLastError());
  const int rc = h->allreduce(h->allreduce_ctx, h->dense_counts.as<u32>(), dense_n, s);
  if (rc)
    return fail(h, SW_ERR_COLLECTIVE, "the caller's all-reduce hook failed");
  launch_node_counts_from_dense(in_key, h->node_start.as<u32>(), n_nodes, shift_for_levels(levels),
                                h->dense_counts.as<u32>(), h->node_gcount.as<u32>(), s);
  h->stats.kernel_launches += 2;
  return SW_OK;
}

int
read_store_vals(swgpu_tiler* h, const u64* a, const u64* b)
{
  CK(cudaMemcpyAsync(&h->h_scalars->store_vals[0], a, 8, cudaMemcpyDeviceToHost, h->stream));
  if (b)
    CK(cudaMemcpyAsync(&h->h_scalars->store_vals[1], b, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SW_OK;
}

// does the selection of this node level run the minimum-distance greedy (sweep_level's switch)
bool
level_uses_min_distance(const swgpu_tiler* h, int node_level)
{
  if (h->prm.sampling == SW_MIN_DISTANCE)
    return true;
  if (h->prm.sampling == SW_MIN_DISTANCE_FAST)
    return cand_level_sampler(h, node_level) >= 0;
  return false;
}

// MIN_DISTANCE on a node that spans GPUs: accepted points near a shard face are exchanged, the point of the
// higher rank loses a conflict (kernels_shard.cu).  Collective: every rank calls it once per spanning level
// that is sampled, m == nullptr when it has no points there.
int
resolve_shard_faces(swgpu_tiler* h, const SwMinDistArgs* m)
{
  cudaStream_t s = h->stream;
  u64 n_mine = 0;
  const u64 count = m ? m->count : 0;
  unsigned char* state = static_cast<unsigned char*>(h->md.state.p);
  if (count) {
    CK(h->face_flags.ensure(count * 4));
    CK(h->face_offs.ensure((count + 1) * 8));
    CK(h->face_scan.ensure(scan_scratch_words(count) * 8));
    const double reach = std::sqrt(m->threshold) * (1.0 + 1e-6);
    launch_face_flag(m->in_idx, m->pos_sorted, state, count, h->bounds, reach, h->face_first_prefix, h->face_n_ranks,
                     h->face_rank, h->face_flags.as<u32>(), s);
    launch_exclusive_scan_u32(h->face_flags.as<u32>(), count, h->face_offs.as<u64>(), h->face_scan.as<u64>(), s);
    h->stats.kernel_launches += 4;
    CK(cudaGetLastError());
    const int rc = read_store_vals(h, h->face_offs.as<u64>() + count, nullptr);
    if (rc)
      return rc;
    n_mine = h->h_scalars->store_vals[0];
    CK(h->face_rec.ensure(std::max<u64>(n_mine, 1) * sizeof(SwFaceRecord)));
    CK(h->face_src.ensure(std::max<u64>(n_mine, 1) * 4));
    launch_face_collect(m->in_key, m->in_idx, m->pos_sorted, count, h->face_flags.as<u32>(), h->face_offs.as<u64>(),
                        h->face_rec.as<SwFaceRecord>(), h->face_src.as<u32>(), s);
    h->stats.kernel_launches += 1;
    CK(cudaGetLastError());
  } else {
    CK(h->face_rec.ensure(sizeof(SwFaceRecord)));
  }
  void* all = nullptr;
  uint64_t bytes[SW_MAX_RANKS] = {};
  if (h->face_fn(h->face_ctx, h->face_rec.p, n_mine * sizeof(SwFaceRecord), &all, bytes, s) != 0)
    return fail(h, SW_ERR_COLLECTIVE, "the caller's all-gather hook failed (shard faces)");
  h->face_points_sent += n_mine;
  if (!n_mine || !h->face_rank)
    return SW_OK;
  SwFaceRanks fr{};
  for (u32 r = 0; r < h->face_n_ranks; ++r)
    fr.first[r + 1] = fr.first[r] + bytes[r] / sizeof(SwFaceRecord);
  if (fr.first[h->face_rank] == 0)
    return SW_OK; // no lower rank has points near a face
  launch_face_resolve(h->face_rec.as<SwFaceRecord>(), h->face_src.as<u32>(), (u32)n_mine,
                      static_cast<const SwFaceRecord*>(all), fr, h->face_rank, m->cell_levels, m->node_levels,
                      m->threshold, state, s);
  h->stats.kernel_launches += 1;
  CK(cudaGetLastError());
  return SW_OK;
}

// One sampling level over the list [in_key, in_idx) of `count` points whose nodes have `levels`
// levels (reference node level = levels - 1).  Appends the selected points to the output arrays
// and, if rem_key != nullptr, writes the remainder list.
int
sweep_level(swgpu_tiler* h, const u64* in_key, const u32* in_idx, u64 count, int levels, bool allow_take_all,
            bool force_all, u64* rem_key, u32* rem_idx, u32 chunk_flags, u64* n_selected_out, bool input_in_out_buffers,
            u64 input_out_offset, bool nodes_known, u32 n_nodes_known, u32* n_nodes_next_out)
{
  const int node_level = levels - 1;
  const int node_shift = shift_for_levels(levels);
  cudaStream_t s = h->stream;

  if (h->multi_batch) { // list lengths are not bounded by the batch size
    const size_t tiles = sweep_tiles(count);
    CK(h->tile_rank0.ensure(tiles * 4));
    CK(h->scan_status.ensure(tiles * 8 * 5));
    if (needs_positions(h->prm.sampling))
      CK(h->sel.ensure(count));
  }
  // grow the append-only outputs before taking pointers into them
  CK(h->out_key.ensure((h->out_count + count) * 8, s, h->out_count * 8));
  CK(h->out_idx.ensure((h->out_count + count) * 4, s, h->out_count * 4));
  if (input_in_out_buffers) {
    in_key = h->out_key.as<u64>() + input_out_offset;
    in_idx = h->out_idx.as<u32>() + input_out_offset;
  }

  // node boundaries: known from the previous level's child counts, else found by a pass over the keys
  int rc = SW_OK;
  u32 n_nodes = n_nodes_known;
  if (nodes_known) {
    launch_tile_rank0(h->node_start.as<u32>(), n_nodes, count, h->tile_rank0.as<u32>(), s);
    h->stats.kernel_launches += 1;
  } else {
    launch_node_rle(in_key, count, node_shift, h->node_start.as<u32>(), h->tile_rank0.as<u32>(), h->d_n_nodes(),
                    h->scan_status.as<u64>(), h->d_tickets(), s);
    h->stats.kernel_launches += 1;
    h->stats.bytes_traffic += 8 * count;
    rc = sync_scalars(h);
    if (rc)
      return rc;
    n_nodes = h->h_scalars->n_nodes;
  }
  CK(h->node_index.ensure((h->node_count + n_nodes) * 8, s, h->node_count * 8));
  CK(h->node_first.ensure((h->node_count + n_nodes + 1) * 8, s, h->node_count * 8));

  // nodes above the shard prefix depth span GPUs: take-all needs their global point count
  const u32* node_gcount = h->gcount_override; // multi-batch: nodes with stored points are never taken whole
  if (allow_take_all && !force_all && spans_shards(h, levels)) {
    rc = exchange_node_counts(h, in_key, n_nodes, levels);
    if (rc)
      return rc;
    node_gcount = h->node_gcount.as<u32>();
  }

  bool reads_positions = false; // does this level's selection read the 24-byte positions
  SwLevelArgs a{};
  a.in_key = in_key;
  a.in_idx = in_idx;
  a.count = count;
  a.node_shift = node_shift;
  a.cell_shift = 63;
  a.sampling = h->prm.sampling;
  a.force_all = force_all ? 1 : 0;
  a.allow_take_all = allow_take_all ? 1 : 0;
  a.max_points_per_node = h->prm.max_points_per_node;
  a.node_start = h->node_start.as<u32>();
  a.tile_rank0 = h->tile_rank0.as<u32>();
  a.node_gcount = node_gcount;
  a.sel = nullptr;

  if (!force_all) {
    const int cand = cand_level_sampler(h, node_level);
    int sampling = h->prm.sampling;
    // GridCenterSampling takes the first point when the candidate level is the root
    // (`return ++partition_point`, Sampling.h:346-348): same selection as RANDOM_GRID at level -1
    // (AdaptivePoissonDiskSampling has the same early return, Sampling.h:513-515)
    if ((sampling == SW_GRID_CENTER || sampling == SW_MIN_DISTANCE_FAST) && cand < 0) {
      sampling = SW_RANDOM_GRID;
      a.sampling = SW_RANDOM_GRID;
    }
    switch (sampling) {
      case SW_RANDOM_GRID: {
        if (cand >= 21)
          return fail(h, SW_ERR_DEEP_REROOT, "sampling grid deeper than MortonIndex64 (re-root path not supported)");
        a.cell_shift = cand < 0 ? 63 : 3 * (20 - cand);
        break;
      }
      case SW_GRID_CENTER:
      case SW_JITTERED: {
        if (h->prm.sampling == SW_GRID_CENTER && cand >= 21)
          return fail(h, SW_ERR_DEEP_REROOT, "sampling grid deeper than MortonIndex64 (re-root path not supported)");
        CK(cudaMemsetAsync(h->sel.p, 0, count, s));
        CK(cudaMemsetAsync(h->d_error(), 0, 4, s));
        SwArgminArgs g{};
        g.in_key = in_key;
        g.in_idx = in_idx;
        g.count = count;
        g.pos_sorted = h->sample_pos;
        g.sampling = h->prm.sampling;
        g.node_shift = node_shift;
        g.node_level = node_level;
        g.cand_level = cand;
        g.cell_shift = cand < 0 ? 63 : 3 * (20 - cand);
        g.spacing_at_node = h->prm.spacing_at_root / std::pow(2, node_level + 1);
        g.bounds = h->bounds;
        g.sel = h->sel.as<unsigned char>();
        g.error_flag = h->d_error();
        g.node_start = h->node_start.as<u32>();
        g.tile_rank0 = h->tile_rank0.as<u32>();
        g.node_gcount = node_gcount;
        g.allow_take_all = allow_take_all ? 1 : 0;
        g.max_points_per_node = h->prm.max_points_per_node;
        CK(h->argmin_nodes.ensure(std::max<size_t>(n_nodes, 1) * sizeof(SwArgminNode)));
        g.n_nodes = n_nodes;
        g.nodes = h->argmin_nodes.as<SwArgminNode>();
        launch_select_argmin(g, h->scan_status.as<u64>(), h->d_tickets(), s);
        h->stats.kernel_launches += 2;
        h->stats.bytes_traffic += (8 + 4 + 24 + 1 + 1) * count;
        reads_positions = true;
        a.sel = h->sel.as<unsigned char>();
        break;
      }
      case SW_MIN_DISTANCE:
      case SW_MIN_DISTANCE_FAST: {
        SwMinDistArgs m{};
        m.in_key = in_key;
        m.in_idx = in_idx;
        m.count = count;
        m.pos_sorted = h->sample_pos;
        m.node_shift = node_shift;
        m.node_levels = levels;
        const double spacing_at_node = h->prm.spacing_at_root / std::pow(2, node_level + 1);
        const float sf = static_cast<float>(spacing_at_node);
        const float sq = sf * sf; // SparseGrid.cpp:11-14: squared in float
        m.threshold = static_cast<double>(sq);
        // Morton cells whose side is >= spacing * (1 + 1e-6): two points closer than the spacing lie
        // in the same or in adjacent cells.  side(levels) = extent / 2^levels.
        {
          const double ratio = root_extent_x(h) / (spacing_at_node * (1.0 + 1e-6));
          int cl = ratio >= 1.0 ? static_cast<int>(std::floor(std::log2(ratio))) : 0;
          if (cl > 21)
            cl = 21;
          if (cl < 0)
            cl = 0;
          m.cell_levels = cl;
        }
        m.node_start = h->node_start.as<u32>();
        m.tile_rank0 = h->tile_rank0.as<u32>();
        m.node_gcount = node_gcount;
        m.allow_take_all = allow_take_all ? 1 : 0;
        m.max_points_per_node = h->prm.max_points_per_node;
        // MIN_DISTANCE_FAST analyses every n-th point of a node only: n = round(1 / density(level)) with
        // the CLI's density function (process/TilerProcess.cpp:500-508; Sampling.h:525-539)
        m.nth_point = 1;
        if (sampling == SW_MIN_DISTANCE_FAST)
          m.nth_point = node_level < 0 ? 4u : (node_level < 1 ? 2u : 1u);
        h->md.status = h->scan_status.as<u64>();
        h->md.ticket = h->d_tickets();
        h->md.h_pinned = h->h_scalars->scratch;
        u32 rounds = 0, launches = 0;
        u64 bytes = 0;
        cudaError_t e = run_min_distance(m, h->md, s, &rounds, &launches, &bytes);
        if (e != cudaSuccess)
          return fail_cuda(h, e, "run_min_distance");
        h->stats.min_distance_rounds += rounds;
        h->stats.kernel_launches += launches;
        h->stats.bytes_traffic += bytes;
        if (spans_shards(h, levels) && h->face_fn) {
          rc = resolve_shard_faces(h, &m);
          if (rc)
            return rc;
        }
        reads_positions = true;
        a.sel = static_cast<const unsigned char*>(h->md.state.p);
        break;
      }
      default:
        return fail(h, SW_ERR_INVALID_ARGUMENT, "unknown sampling strategy");
    }
  }

  a.out_key = h->out_key.as<u64>();
  a.out_idx = h->out_idx.as<u32>();
  a.out_offset = h->out_count;
  a.rem_key = rem_key;
  a.rem_idx = rem_idx;
  a.node_index = h->node_index.as<u64>();
  a.node_first = h->node_first.as<u64>();
  a.node_base = h->node_count;
  a.levels = levels;
  CK(h->selbits.ensure(sweep_tiles(count) * (SW_SWEEP_TILE / 32) * 4));
  CK(h->tile_sel.ensure(sweep_tiles(count) * 4));
  a.selbits = h->selbits.as<u32>();
  a.tile_sel = h->tile_sel.as<u32>();
  a.status = h->two_pass_compaction ? nullptr : h->scan_status.as<u64>(); // >= sweep_tiles(count) u64 (node_rle)
  a.ticket = h->d_tickets() + 7;
  // the points that stay are counted per child node: that IS the node table of the next level
  const bool want_children = rem_key != nullptr && !force_all && levels < 21;
  const u32 n_child_slots = want_children ? 8u * n_nodes : 0u;
  a.child_count = nullptr;
  if (want_children) {
    CK(h->child_count.ensure((size_t)n_child_slots * 4));
    a.child_count = h->child_count.as<u32>();
  }
  launch_level_compact(a, h->d_n_selected(), n_child_slots, h->node_start_next.as<u32>(), h->d_n_nodes_next(), s);
  h->stats.kernel_launches += a.status ? (a.child_count ? 2 : 1) : 3;
  rc = sync_scalars(h);
  if (rc)
    return rc;
  if (h->h_scalars->error_flag)
    return fail(h,
                (int)h->h_scalars->error_flag,
                h->h_scalars->error_flag == SW_ERR_JITTER_GRID_TOO_SMALL
                  ? "Grids smaller than 16x16 are not supported currently!"
                  : "Node is too small to be sampled with ImprovedPoissonSampling!");
  const u64 n_sel = h->h_scalars->n_selected;
  // algorithmic bytes of the level by SURVEY 8(d): read key + id (+ position), write the compacted remainder,
  // write the selected ids
  h->stats.bytes_sample += (12 + (reads_positions ? 24 : 0)) * count + (rem_key ? 12 * (count - n_sel) : 0) + 4 * n_sel;
  // what this design really moves: keys + ids read once (the two-pass kernels read the keys twice), (key, id) written
  h->stats.bytes_traffic += ((a.status ? 8 : 16) + (in_idx ? 4 : 0)) * count + 12 * n_sel + (rem_key ? 12 * (count - n_sel) : 0);
  if (n_nodes_next_out)
    *n_nodes_next_out = want_children ? h->h_scalars->n_nodes_next : 0u;
  if (want_children) // the next level reads its node boundaries from node_start
    std::swap(h->node_start, h->node_start_next);
  h->stats.sweep_points += count;

  Chunk c{};
  c.levels = levels;
  c.out_offset = h->out_count;
  c.count = n_sel;
  c.node_base = h->node_count;
  c.n_nodes = n_nodes;
  c.flags = chunk_flags | (force_all ? SW_NODE_TERMINAL : 0u);
  h->chunks.push_back(c);
  h->out_count += n_sel;
  h->node_count += n_nodes;
  *n_selected_out = n_sel;
  return SW_OK;
}

// estimate_start_node_level_in_octree, TilingAlgorithms.cpp:1473-1535.  `count_of(b, group)` returns
// the number of points in the level-5 prefixes [b, b + group).
template<typename CountOf>
int
start_level_from_prefix_counts(CountOf count_of, size_t concurrency)
{
  const u32 MIN_LEVEL = 3, MAX_LEVEL = 6;
  for (u32 level = 0; level < MAX_LEVEL; ++level) {
    // ranges at `level` = prefixes of (level + 1) levels = groups of 8^(5 - level) bins
    const u32 group = 1u << (3 * (5 - level));
    size_t ranges = 0, large = 0;
    for (u32 b = 0; b < 262144u; b += group) {
      const u64 cnt = count_of(b, group);
      if (cnt > 0) {
        ++ranges;
        if (cnt >= 100000u)
          ++large;
      }
    }
    float score = 0.f;
    if (!(ranges <= concurrency / 2))
      score = static_cast<float>(large) / static_cast<float>(concurrency);
    if (score >= 1.f)
      return (int)std::max(level + 1, MIN_LEVEL);
  }
  return (int)MAX_LEVEL;
}

// single GPU: bin boundaries by binary search in the sorted keys
int
estimate_start_level(swgpu_tiler* h, int* S_out)
{
  launch_level5_bins(h->keys[0].as<u64>(), h->n, h->bins.as<u32>(), h->stream);
  h->stats.kernel_launches += 1;
  std::vector<u32> bins(262145);
  CK(cudaMemcpyAsync(bins.data(), h->bins.p, 262145 * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *S_out = start_level_from_prefix_counts([&](u32 b, u32 group) -> u64 { return bins[b + group] - bins[b]; },
                                          h->prm.concurrency);
  return SW_OK;
}

// one shard of a multi-GPU run: the same estimate on GLOBAL counts.  Every rank takes its level-5 bin
// counts from its sorted keys, the caller's all-reduce hook sums them (1 MB), every rank evaluates the
// same rule on the same numbers.  Collective: called by every rank, empty shards included.
int
estimate_start_level_global(swgpu_tiler* h, int* S_out)
{
  launch_level5_bins(h->keys[0].as<u64>(), h->n, h->bins.as<u32>(), h->stream);
  CK(h->dense_counts.ensure(262144 * 4));
  launch_bin_counts(h->bins.as<u32>(), 262144u, h->dense_counts.as<u32>(), h->stream);
  h->stats.kernel_launches += 2;
  CK(cudaGetLastError());
  if (h->allreduce(h->allreduce_ctx, h->dense_counts.as<u32>(), 262144u, h->stream) != 0)
    return fail(h, SW_ERR_COLLECTIVE, "the all-reduce hook failed (start level)");
  std::vector<u32> counts(262144);
  CK(cudaMemcpyAsync(counts.data(), h->dense_counts.p, 262144 * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  std::vector<u64> prefix(262145, 0);
  for (u32 b = 0; b < 262144u; ++b)
    prefix[b + 1] = prefix[b] + counts[b];
  *S_out = start_level_from_prefix_counts([&](u32 b, u32 group) -> u64 { return prefix[b + group] - prefix[b]; },
                                          h->prm.concurrency);
  return SW_OK;
}

int
run_batch(swgpu_tiler* h)
{
  const u64 n = h->n;
  cudaStream_t s = h->stream;
  h->chunks.clear();
  h->out_count = 0;
  h->node_count = 0;
  h->start_level = -1;
  h->batch_done = false;
  h->finalized = false;
  h->n_clamped = 0;
  h->h_scalars->n_clamped = 0; // an empty shard never refreshes the pinned copy
  std::memset(&h->stats, 0, sizeof(h->stats));
  h->stats.n_points = n;

  if (n >= (1ull << 30))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "a batch is limited to 2^30 - 1 points per GPU");
  // The level tables, the MIN_DISTANCE cell size and the jitter grid are derived from the x extent, as in the
  // reference, which always tiles against AABB::makeCubic bounds (process/Tiler.cpp:185-187).  Non-cubic bounds
  // would silently break the 27-neighbour search, so tiling refuses them (extents may differ by rounding only);
  // the stand-alone indexing / sort primitives accept any box, like index_point.
  {
    const double ex = h->prm.bounds_max[0] - h->prm.bounds_min[0];
    for (int a = 1; a < 3; ++a) {
      const double e = h->prm.bounds_max[a] - h->prm.bounds_min[a];
      if (std::fabs(e - ex) > 1e-9 * std::fabs(ex))
        return fail(h, SW_ERR_INVALID_ARGUMENT, "tiling needs cubic bounds (AABB::makeCubic)");
    }
  }
  const bool sharded = h->shard_levels > 0; // the caller checked the GLOBAL point count
  if (!sharded && h->prm.tiling == SW_FAST && n < h->prm.concurrency)
    return fail(h, SW_ERR_TOO_FEW_POINTS, "Can't scatter a range that has less than 'scatter_factor' elements!");
  if (!sharded && n == 0)
    return fail(h, SW_ERR_EMPTY_NODE, "tile_internal_node: Got zero points to tile @ node r");
  int rc = ensure_batch_buffers(h, n);
  if (rc)
    return rc;

  record(h, 0);
  // K1 (every launcher is a no-op for an empty shard)
  CK(cudaMemsetAsync(h->hist.p, 0, sort_hist_words() * 4, s));
  CK(cudaMemsetAsync(h->scalars.p, 0, 128, s));
  h->sort_first_pass = choose_sort_first_pass(h);
  u64* unsorted_keys = h->keys[sort_input_buffer_top(h->sort_first_pass)].as<u64>(); // the last pass lands in keys[0]
  if (h->d_las) { // K1-LAS: 12 B record in, 24 B position + 8 B key out
    launch_las_encode(h->d_las, n, h->las_t, h->bounds, h->d_xyz, unsorted_keys, h->hist.as<u32>(),
                      h->d_n_clamped(), s);
    h->stats.bytes_index = 48 * n; // 12 B record + 24 B position + 8 B key + 4 B id
    h->stats.bytes_traffic += 44 * n; // the ids are generated by the first sort pass, never written here
  } else {
    launch_morton_encode(h->d_xyz, n, h->bounds, unsorted_keys, h->hist.as<u32>(), h->d_n_clamped(), s);
    h->stats.bytes_index = 36 * n; // 24 B position + 8 B key + 4 B id
    h->stats.bytes_traffic += 32 * n;
  }
  h->stats.kernel_launches += 1;
  record(h, 1);
  // K2
  rc = sort_batch(h, n);
  if (rc)
    return rc;
  record(h, 2);
  // K4
  if (needs_positions(h->prm.sampling)) {
    launch_gather_positions(h->d_xyz, h->vals[0].as<u32>(), n, h->pos_sorted.as<double>(), s);
    h->stats.kernel_launches += 1;
    h->stats.bytes_gather = 2 * 24 * n; // SURVEY 8(d) K4 without attribute bytes
    h->stats.bytes_traffic += (4 + 24 + 24) * n;
  }
  record(h, 3);
  CK(cudaGetLastError());

  h->sample_pos = h->pos_sorted.as<double>();
  h->gcount_override = nullptr;
  int first_levels = 0; // ACCURATE: root has 0 levels
  if (h->prm.tiling == SW_FAST) {
    int S = 0;
    if (sharded && h->start_level_override >= 0) {
      S = h->start_level_override; // estimated by the caller from the all-reduced prefix histogram
      launch_level5_bins(h->keys[0].as<u64>(), h->n, h->bins.as<u32>(), h->stream); // start nodes below
      h->stats.kernel_launches += 1;
    } else {
      rc = (sharded && h->allreduce) ? estimate_start_level_global(h, &S) : estimate_start_level(h, &S);
      if (rc)
        return rc;
    }
    h->start_level = S;
    first_levels = S;
  }

  // level-synchronous sweep
  const u64* in_key = h->keys[0].as<u64>();
  h->out_idx_original = !needs_positions(h->prm.sampling);
  const u32* in_idx = h->out_idx_original ? h->vals[0].as<u32>() : nullptr;
  u64 count = n;
  u64* rem_key[2] = { h->keys[1].as<u64>(), h->wkey2.as<u64>() };
  u32* rem_idx[2] = { h->vals[1].as<u32>(), h->widx2.as<u32>() };
  int flip = 0;
  // node boundaries of the first level: the root is one node; FAST's start nodes are found by one
  // run-length pass over the sorted keys; every later level gets them from the child counts
  bool nodes_known = false;
  u32 n_nodes_known = 0;
  if (first_levels == 0 && count > 0) {
    launch_root_node(h->node_start.as<u32>(), count, s);
    nodes_known = true;
    n_nodes_known = 1;
  } else if (count > 0) {
    launch_start_nodes(h->bins.as<u32>(), first_levels, count, h->node_start.as<u32>(), h->d_n_nodes(), s);
    h->stats.kernel_launches += 1;
    rc = sync_scalars(h);
    if (rc)
      return rc;
    nodes_known = true;
    n_nodes_known = h->h_scalars->n_nodes;
  }
  for (int levels = first_levels; count > 0 || spans_shards(h, levels); ++levels) {
    const int node_level = levels - 1;
    const LevelKind kind = level_kind(h, node_level);
    const bool deep = kind == KIND_REROOT && h->deep_policy == 1 && levels <= 21;
    if ((kind == KIND_REROOT && !deep) || levels > 21)
      return fail(h, SW_ERR_DEEP_REROOT, "deep re-root path (TilingAlgorithms.cpp:444-483) is not supported");
    const bool terminal = (kind == KIND_TERMINAL) || deep;
    if (!terminal && levels >= 21)
      return fail(h, SW_ERR_DEEP_REROOT, "child level exceeds MortonIndex64 capacity");
    if (count == 0) { // nothing left on this GPU, but the other shards still need our (zero) counts
      if (!terminal) {
        rc = exchange_node_counts(h, in_key, 0, levels);
        if (rc)
          return rc;
        if (h->face_fn && level_uses_min_distance(h, node_level)) {
          rc = resolve_shard_faces(h, nullptr);
          if (rc)
            return rc;
        }
      }
      continue;
    }
    u64 n_sel = 0;
    u32 n_nodes_next = 0;
    rc = sweep_level(h, in_key, in_idx, count, levels, /*allow_take_all=*/true, terminal, rem_key[flip],
                     rem_idx[flip], deep ? SW_NODE_DEEP : 0u, &n_sel, false, 0, nodes_known, n_nodes_known,
                     &n_nodes_next);
    if (rc)
      return rc;
    h->stats.n_levels += 1;
    in_key = rem_key[flip];
    in_idx = rem_idx[flip];
    flip ^= 1;
    count -= n_sel;
    nodes_known = true;
    n_nodes_known = n_nodes_next;
  }
  record(h, 4);
  CK(cudaGetLastError());
  h->n_clamped = h->h_scalars->n_clamped;
  h->batch_done = true;
  return SW_OK;
}

int
run_finalize(swgpu_tiler* h)
{
  if (!h->batch_done)
    return SW_OK; // build_execution_graph never ran (TilingAlgorithms.cpp:1242-1246)
  if (h->prm.tiling != SW_FAST || h->finalized)
    return SW_OK;
  const int S = h->start_level;
  if (h->chunks.empty()) { // empty shard: only the collectives of the spanning levels
    for (int lv = S - 1; lv >= 0; --lv)
      if (spans_shards(h, lv) && h->face_fn && level_uses_min_distance(h, lv - 1)) {
        const int rc = resolve_shard_faces(h, nullptr);
        if (rc)
          return rc;
      }
    h->finalized = true;
    return SW_OK;
  }
  // input of the first reconstruct level: the chunk of the start nodes themselves
  size_t src = 0; // chunks[0] has levels == S
  for (int lv = S - 1; lv >= 0; --lv) {
    const Chunk in = h->chunks[src];
    u64 n_sel = 0;
    launch_parent_nodes(h->node_index.as<u64>() + in.node_base, h->node_first.as<u64>() + in.node_base, in.n_nodes,
                        in.out_offset, in.count, h->node_start.as<u32>(), h->d_n_nodes(), h->stream);
    h->stats.kernel_launches += 1;
    int rc = sync_scalars(h);
    if (rc)
      return rc;
    const u32 n_parents = h->h_scalars->n_nodes;
    rc = sweep_level(h, nullptr, nullptr, in.count, lv, /*allow_take_all=*/false, false, nullptr, nullptr,
                     SW_NODE_RECONSTRUCTED, &n_sel, true, in.out_offset, true, n_parents, nullptr);
    if (rc)
      return rc;
    h->stats.n_reconstruct_levels += 1;
    src = h->chunks.size() - 1;
  }
  record(h, 5);
  h->finalized = true;
  return SW_OK;
}


// =============================================================================================
// multi-batch mode: TilingAlgorithmV1 / V3 over several batches against a node store in HBM
// =============================================================================================
// The selection of one sweep level (the last chunk) replaces what the visited nodes stored; every other node of
// the level keeps its points.  Table and id pool are rebuilt in node-index order.
int
store_update(swgpu_tiler* h, int levels, const Chunk& c)
{
  LevelStore& L = h->store[levels];
  cudaStream_t s = h->stream;
  const u32 nv = c.n_nodes, no = (u32)L.n_nodes;
  if (!nv)
    return SW_OK;
  const u64* vidx = h->node_index.as<u64>() + c.node_base;
  const u64* vfirst = h->node_first.as<u64>() + c.node_base;
  const size_t nn_max = (size_t)nv + no;
  CK(h->st_lo.ensure((size_t)nv * 4));
  CK(h->st_found.ensure((size_t)nv * 4));
  CK(h->st_cumf.ensure(((size_t)nv + 1) * 8));
  CK(h->st_scan.ensure(scan_scratch_words(nn_max) * 8));
  CK(ensure_amortised(L.spare_index, nn_max * 8));
  CK(h->st_ncnt.ensure(nn_max * 4));
  CK(ensure_amortised(L.spare_flags, nn_max * 4));
  CK(h->st_nsrc.ensure(nn_max * 8));
  CK(ensure_amortised(L.spare_first, (nn_max + 1) * 8));
  launch_store_match(vidx, nv, L.index.as<u64>(), no, h->st_lo.as<u32>(), h->st_found.as<u32>(), s);
  launch_exclusive_scan_u32(h->st_found.as<u32>(), nv, h->st_cumf.as<u64>(), h->st_scan.as<u64>(), s);
  CK(cudaMemsetAsync(h->st_ncnt.p, 0, nn_max * 4, s)); // rows past the real table count nothing
  launch_store_rows(vidx, vfirst, nv, c.out_offset + c.count, c.flags, h->st_lo.as<u32>(), h->st_cumf.as<u64>(),
                    L.index.as<u64>(), L.first.as<u64>(), L.flags.as<u32>(), no, L.spare_index.as<u64>(),
                    h->st_ncnt.as<u32>(), L.spare_flags.as<u32>(), h->st_nsrc.as<u64>(), s);
  launch_exclusive_scan_u32(h->st_ncnt.as<u32>(), nn_max, L.spare_first.as<u64>(), h->st_scan.as<u64>(), s);
  h->stats.kernel_launches += 8;
  CK(cudaGetLastError());
  int rc = read_store_vals(h, h->st_cumf.as<u64>() + nv, L.spare_first.as<u64>() + nn_max);
  if (rc)
    return rc;
  const u64 n_found = h->h_scalars->store_vals[0], total = h->h_scalars->store_vals[1];
  const u64 nn = nn_max - n_found;
  CK(ensure_amortised(L.spare_ids, std::max<u64>(total, 1) * 4));
  launch_store_copy(total, L.spare_first.as<u64>(), (u32)nn, h->st_nsrc.as<u64>(), L.ids.as<u32>(),
                    h->out_idx.as<u32>(), L.spare_ids.as<u32>(), s);
  h->stats.kernel_launches += 1;
  h->stats.bytes_traffic += 8 * total;
  CK(cudaGetLastError());
  std::swap(L.index, L.spare_index);
  std::swap(L.first, L.spare_first);
  std::swap(L.flags, L.spare_flags);
  std::swap(L.ids, L.spare_ids);
  L.n_nodes = nn;
  L.n_ids = total;
  return SW_OK;
}

// One sweep level of a batch against the store.  cur: slot of list_key / list_idx that holds the input list
// (Morton ordered, node boundaries in h->node_start); on return the slot of the remainder list.
int
store_level(swgpu_tiler* h, int* cur, u64* count_io, int levels, bool terminal, u32 level_flags, u32 n_nodes,
            u32* n_nodes_next)
{
  cudaStream_t s = h->stream;
  LevelStore& L = h->store[levels];
  int free_slots[3], nf = 0;
  for (int i = 0; i < 4; ++i)
    if (i != *cur)
      free_slots[nf++] = i;
  const int sb = free_slots[0], sc = free_slots[1];
  int sr = free_slots[2];
  u64 count = *count_io;
  const u64* in_key = h->list_key[*cur].as<u64>();
  const u32* in_idx = h->list_idx[*cur].as<u32>();
  int in_slot = *cur;

  u64 m = 0;
  if (L.n_nodes) { // which of the visited nodes hold points from earlier batches
    CK(h->st_slot.ensure((size_t)n_nodes * 4));
    CK(h->st_cnt.ensure((size_t)n_nodes * 4));
    CK(h->st_boff.ensure(((size_t)n_nodes + 1) * 8));
    CK(h->st_scan.ensure(scan_scratch_words(n_nodes) * 8));
    launch_store_lookup(in_key, h->node_start.as<u32>(), n_nodes, shift_for_levels(levels), L.index.as<u64>(),
                        (u32)L.n_nodes, L.first.as<u64>(), h->st_slot.as<u32>(), h->st_cnt.as<u32>(), s);
    launch_exclusive_scan_u32(h->st_cnt.as<u32>(), n_nodes, h->st_boff.as<u64>(), h->st_scan.as<u64>(), s);
    h->stats.kernel_launches += 4;
    CK(cudaGetLastError());
    const int rc = read_store_vals(h, h->st_boff.as<u64>() + n_nodes, nullptr);
    if (rc)
      return rc;
    m = h->h_scalars->store_vals[0];
  }
  h->gcount_override = nullptr;
  if (m) {
    if (count + m >= (1ull << 30))
      return fail(h, SW_ERR_INVALID_ARGUMENT, "a sweep list is limited to 2^30 - 1 points (batch + revisited nodes)");
    CK(h->list_key[sb].ensure(m * 8));
    CK(h->list_idx[sb].ensure(m * 4));
    CK(h->list_key[sc].ensure((count + m) * 8));
    CK(h->list_idx[sc].ensure((count + m) * 4));
    CK(h->node_start_next.ensure((count + m + 1) * 4));
    CK(h->st_gcount.ensure((size_t)n_nodes * 4));
    launch_store_fetch(m, h->st_boff.as<u64>(), n_nodes, h->st_slot.as<u32>(), L.first.as<u64>(), L.ids.as<u32>(),
                       in_key, h->node_start.as<u32>(), levels, h->store_xyz.as<double>(), h->bounds,
                       h->list_key[sb].as<u64>(), h->list_idx[sb].as<u32>(), s);
    if (terminal) // merge_node_data_unsorted, Node.cpp:22-34
      launch_concat_lists(in_key, in_idx, count, h->list_key[sb].as<u64>(), h->list_idx[sb].as<u32>(), m,
                          h->node_start.as<u32>(), h->st_boff.as<u64>(), n_nodes, h->list_key[sc].as<u64>(),
                          h->list_idx[sc].as<u32>(), s);
    else // merge_node_data_sorted, Node.cpp:3-20
      launch_merge_lists(in_key, in_idx, count, h->list_key[sb].as<u64>(), h->list_idx[sb].as<u32>(), m,
                         h->list_key[sc].as<u64>(), h->list_idx[sc].as<u32>(), s);
    launch_store_merged_nodes(h->node_start.as<u32>(), h->st_boff.as<u64>(), n_nodes, h->node_start_next.as<u32>(),
                              h->st_gcount.as<u32>(), s);
    std::swap(h->node_start, h->node_start_next);
    h->stats.kernel_launches += 3;
    h->stats.bytes_traffic += (24 + 12 + 4) * m + 24 * (count + m);
    CK(cudaGetLastError());
    count += m;
    in_key = h->list_key[sc].as<u64>();
    in_idx = h->list_idx[sc].as<u32>();
    in_slot = sc;
    h->gcount_override = h->st_gcount.as<u32>();
  }
  if (sr == in_slot)
    sr = *cur;
  CK(h->list_key[sr].ensure(count * 8));
  CK(h->list_idx[sr].ensure(count * 4));
  CK(h->node_start_next.ensure((count + 1) * 4));

  u64 n_sel = 0;
  int rc = sweep_level(h, in_key, in_idx, count, levels, /*allow_take_all=*/true, terminal, h->list_key[sr].as<u64>(),
                       h->list_idx[sr].as<u32>(), level_flags, &n_sel, false, 0, true, n_nodes, n_nodes_next);
  h->gcount_override = nullptr;
  if (rc)
    return rc;
  rc = store_update(h, levels, h->chunks.back());
  if (rc)
    return rc;
  // the output arrays were scratch for this level only
  h->chunks.clear();
  h->out_count = 0;
  h->node_count = 0;
  *cur = sr;
  *count_io = count - n_sel;
  return SW_OK;
}

int
run_batch_store(swgpu_tiler* h)
{
  const u64 n = h->n;
  cudaStream_t s = h->stream;
  h->chunks.clear();
  h->out_count = 0;
  h->node_count = 0;
  h->batch_done = false;
  h->finalized = false;
  h->n_clamped = 0;
  h->h_scalars->n_clamped = 0;
  std::memset(&h->stats, 0, sizeof(h->stats));
  h->stats.n_points = n;
  if (h->shard_levels)
    return fail(h, SW_ERR_STATE, "multi-batch mode and sharding cannot be combined");
  if (n >= (1ull << 30))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "a batch is limited to 2^30 - 1 points per GPU");
  if (h->store_points + n >= (1ull << 32))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "global point ids are 32 bit: all batches together must stay below 2^32");
  {
    const double ex = h->prm.bounds_max[0] - h->prm.bounds_min[0];
    for (int a = 1; a < 3; ++a) {
      const double e = h->prm.bounds_max[a] - h->prm.bounds_min[a];
      if (std::fabs(e - ex) > 1e-9 * std::fabs(ex))
        return fail(h, SW_ERR_INVALID_ARGUMENT, "tiling needs cubic bounds (AABB::makeCubic)");
    }
  }
  if (h->prm.tiling == SW_FAST && n < h->prm.concurrency)
    return fail(h, SW_ERR_TOO_FEW_POINTS, "Can't scatter a range that has less than 'scatter_factor' elements!");
  if (n == 0)
    return fail(h, SW_ERR_EMPTY_NODE, "tile_internal_node: Got zero points to tile @ node r");
  int rc = ensure_batch_buffers(h, n);
  if (rc)
    return rc;

  record(h, 0);
  CK(cudaMemsetAsync(h->hist.p, 0, sort_hist_words() * 4, s));
  CK(cudaMemsetAsync(h->scalars.p, 0, 128, s));
  h->sort_first_pass = choose_sort_first_pass(h);
  u64* unsorted_keys = h->keys[sort_input_buffer_top(h->sort_first_pass)].as<u64>();
  if (h->d_las) {
    launch_las_encode(h->d_las, n, h->las_t, h->bounds, h->d_xyz, unsorted_keys, h->hist.as<u32>(), h->d_n_clamped(),
                      s);
    h->stats.bytes_index = 48 * n;
  } else {
    launch_morton_encode(h->d_xyz, n, h->bounds, unsorted_keys, h->hist.as<u32>(), h->d_n_clamped(), s);
    h->stats.bytes_index = 36 * n;
  }
  record(h, 1);
  rc = sort_batch(h, n);
  if (rc)
    return rc;
  h->stats.kernel_launches += 1;
  record(h, 2);
  // the batch joins the point store (clamped positions, like the PointBuffer after index_point)
  const u64 base = h->store_points;
  CK(h->store_xyz.ensure((base + n) * 24, s, base * 24));
  CK(cudaMemcpyAsync(h->store_xyz.as<double>() + 3 * base, h->d_xyz, n * 24, cudaMemcpyDeviceToDevice, s));
  h->store_points = base + n;
  h->sample_pos = h->store_xyz.as<double>();
  int cur = 0;
  CK(h->list_key[0].ensure(n * 8));
  CK(h->list_idx[0].ensure(n * 4));
  CK(cudaMemcpyAsync(h->list_key[0].p, h->keys[0].p, n * 8, cudaMemcpyDeviceToDevice, s));
  launch_make_gids(h->vals[0].as<u32>(), n, (u32)base, h->list_idx[0].as<u32>(), s);
  h->stats.kernel_launches += 1;
  record(h, 3);
  CK(cudaGetLastError());

  int first_levels = 0;
  if (h->prm.tiling == SW_FAST) {
    if (h->store_start_level < 0) { // the first batch fixes the start level (TilingAlgorithms.cpp:1250-1360)
      int S = 0;
      rc = estimate_start_level(h, &S);
      if (rc)
        return rc;
      h->store_start_level = S;
    } else { // later batches are cut at it (:1362-1453)
      launch_level5_bins(h->keys[0].as<u64>(), n, h->bins.as<u32>(), s);
      h->stats.kernel_launches += 1;
    }
    h->start_level = h->store_start_level;
    first_levels = h->store_start_level;
  }
  u64 count = n;
  u32 n_nodes = 0;
  if (first_levels == 0) {
    launch_root_node(h->node_start.as<u32>(), count, s);
    n_nodes = 1;
  } else {
    launch_start_nodes(h->bins.as<u32>(), first_levels, count, h->node_start.as<u32>(), h->d_n_nodes(), s);
    h->stats.kernel_launches += 1;
    rc = sync_scalars(h);
    if (rc)
      return rc;
    n_nodes = h->h_scalars->n_nodes;
  }
  for (int levels = first_levels; count > 0; ++levels) {
    const LevelKind kind = level_kind(h, levels - 1);
    const bool deep = kind == KIND_REROOT && h->deep_policy == 1 && levels <= 21;
    if ((kind == KIND_REROOT && !deep) || levels > 21)
      return fail(h, SW_ERR_DEEP_REROOT, "deep re-root path (TilingAlgorithms.cpp:444-483) is not supported");
    const bool terminal = (kind == KIND_TERMINAL) || deep;
    if (!terminal && levels >= 21)
      return fail(h, SW_ERR_DEEP_REROOT, "child level exceeds MortonIndex64 capacity");
    u32 n_nodes_next = 0;
    rc = store_level(h, &cur, &count, levels, terminal, deep ? SW_NODE_DEEP : 0u, n_nodes, &n_nodes_next);
    if (rc)
      return rc;
    h->stats.n_levels += 1;
    n_nodes = n_nodes_next;
  }
  record(h, 4);
  CK(cudaGetLastError());
  rc = sync_scalars(h);
  if (rc)
    return rc;
  h->n_clamped = h->h_scalars->n_clamped;
  h->batch_done = true;
  return SW_OK;
}

// FAST finalize over the store: reconstruct_left_out_nodes for every start node the store holds
// (TilingAlgorithms.cpp:1717-1784); the input of level lv is the whole pool of level lv + 1 in node order.
int
run_finalize_store(swgpu_tiler* h)
{
  if (!h->batch_done || h->prm.tiling != SW_FAST || h->finalized)
    return SW_OK;
  cudaStream_t s = h->stream;
  const int S = h->store_start_level;
  h->sample_pos = h->store_xyz.as<double>();
  h->gcount_override = nullptr;
  for (int lv = S - 1; lv >= 0; --lv) {
    LevelStore& C = h->store[lv + 1];
    if (!C.n_nodes)
      continue;
    const u64 count = C.n_ids;
    if (count >= (1ull << 30))
      return fail(h, SW_ERR_INVALID_ARGUMENT, "a sweep list is limited to 2^30 - 1 points");
    CK(h->list_key[0].ensure(count * 8));
    CK(h->node_start.ensure((std::max<u64>(count, C.n_nodes) + 1) * 4));
    CK(h->node_start_next.ensure((count + 1) * 4));
    launch_store_root_keys(C.ids.as<u32>(), count, h->store_xyz.as<double>(), h->bounds, h->list_key[0].as<u64>(), s);
    launch_parent_nodes(C.index.as<u64>(), C.first.as<u64>(), (u32)C.n_nodes, 0, count, h->node_start.as<u32>(),
                        h->d_n_nodes(), s);
    h->stats.kernel_launches += 2;
    int rc = sync_scalars(h);
    if (rc)
      return rc;
    const u32 n_parents = h->h_scalars->n_nodes;
    h->store[lv].n_nodes = 0; // levels above the start level only ever hold reconstructed nodes
    h->store[lv].n_ids = 0;
    u64 n_sel = 0;
    rc = sweep_level(h, h->list_key[0].as<u64>(), C.ids.as<u32>(), count, lv, /*allow_take_all=*/false, false, nullptr,
                     nullptr, SW_NODE_RECONSTRUCTED, &n_sel, false, 0, true, n_parents, nullptr);
    if (rc)
      return rc;
    rc = store_update(h, lv, h->chunks.back());
    if (rc)
      return rc;
    h->chunks.clear();
    h->out_count = 0;
    h->node_count = 0;
    h->stats.n_reconstruct_levels += 1;
  }
  record(h, 5);
  h->finalized = true;
  return SW_OK;
}

u64
store_node_total(const swgpu_tiler* h)
{
  u64 t = 0;
  for (const LevelStore& L : h->store)
    t += L.n_nodes;
  return t;
}

u64
store_id_total(const swgpu_tiler* h)
{
  u64 t = 0;
  for (const LevelStore& L : h->store)
    t += L.n_ids;
  return t;
}

// final content of the store: nodes by (levels, index), ids node-major
int
store_get_nodes(swgpu_tiler* h, sw_node* nodes, u32* ids, bool ids_on_device)
{
  u64 row = 0, off = 0;
  for (int lv = 0; lv < 22; ++lv) {
    const LevelStore& L = h->store[lv];
    if (!L.n_nodes)
      continue;
    if (nodes) {
      std::vector<u64> index(L.n_nodes), first(L.n_nodes + 1);
      std::vector<u32> flags(L.n_nodes);
      CK(cudaMemcpyAsync(index.data(), L.index.p, L.n_nodes * 8, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(first.data(), L.first.p, (L.n_nodes + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(flags.data(), L.flags.p, L.n_nodes * 4, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      for (u64 k = 0; k < L.n_nodes; ++k) {
        sw_node& nd = nodes[row + k];
        nd.index = index[k];
        nd.levels = (uint32_t)lv;
        nd.flags = flags[k];
        nd.first = off + first[k];
        nd.count = first[k + 1] - first[k];
      }
    }
    if (ids && L.n_ids)
      CK(cudaMemcpyAsync(ids + off, L.ids.p, L.n_ids * 4, ids_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                         h->stream));
    row += L.n_nodes;
    off += L.n_ids;
  }
  CK(cudaStreamSynchronize(h->stream));
  return SW_OK;
}

int
tile_current_batch(swgpu_tiler* h)
{
  return h->multi_batch ? run_batch_store(h) : run_batch(h);
}

} // namespace

// =============================================================================================
// C ABI
// =============================================================================================
static void
release_store(swgpu_tiler* h)
{
  for (LevelStore& L : h->store) {
    L.index.release();
    L.first.release();
    L.flags.release();
    L.ids.release();
    L.spare_index.release();
    L.spare_first.release();
    L.spare_flags.release();
    L.spare_ids.release();
    L.n_nodes = 0;
    L.n_ids = 0;
  }
  h->store_xyz.release();
  h->store_points = 0;
  h->store_start_level = -1;
}

extern "C" {

int
swgpu_create(const sw_params* params, int device, swgpu_handle* out)
{
  if (!params || !out)
    return SW_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (params->sampling < 0 || params->sampling > SW_MIN_DISTANCE_FAST || params->tiling < 0 || params->tiling > 1)
    return SW_ERR_INVALID_ARGUMENT;
  for (int a = 0; a < 3; ++a)
    if (!(params->bounds_max[a] > params->bounds_min[a]))
      return SW_ERR_INVALID_ARGUMENT;
  if (!(params->spacing_at_root > 0.f))
    return SW_ERR_INVALID_ARGUMENT;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return SW_ERR_CUDA; // no CPU fallback
  }
  if (cudaSetDevice(device) != cudaSuccess)
    return SW_ERR_CUDA;
  // Tiler::Tiler refuses spacings that are too small for the bounds (process/Tiler.cpp:178-183)
  const float ratio = std::log2f(static_cast<float>((params->bounds_max[0] - params->bounds_min[0]) /
                                                    params->spacing_at_root));
  if (ratio >= 21.f)
    return SW_ERR_INVALID_ARGUMENT;

  auto* h = new swgpu_tiler();
  if (const char* e = std::getenv("SWGPU_COMPACT"))
    h->two_pass_compaction = std::strcmp(e, "1pass") != 0;
  h->prm = *params;
  if (h->prm.concurrency == 0)
    h->prm.concurrency = 1;
  h->device = device;
  for (int a = 0; a < 3; ++a) {
    h->bounds.min[a] = params->bounds_min[a];
    h->bounds.max[a] = params->bounds_max[a];
    // std::pow(2, MaxLevels) / node_bounds.extent(), OctreeAlgorithms.h:69
    h->bounds.scale[a] = 2097152.0 / (params->bounds_max[a] - params->bounds_min[a]);
  }
  if (cudaMallocHost(reinterpret_cast<void**>(&h->h_scalars), sizeof(HostScalars)) != cudaSuccess ||
      h->scalars.ensure(256) != cudaSuccess) {
    delete h;
    return SW_ERR_CUDA;
  }
  std::memset(h->h_scalars, 0, sizeof(HostScalars));
  for (auto& e : h->ev)
    cudaEventCreate(&e);
  *out = h;
  return SW_OK;
}

void
swgpu_destroy(swgpu_handle h)
{
  if (!h)
    return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  DevBuf* bufs[] = { &h->las_own, &h->payload_tmp, &h->node_hdr, &h->argmin_nodes, &h->xyz_own,    &h->keys[0],   &h->keys[1],       &h->vals[0],     &h->vals[1],  &h->wkey2,
                     &h->widx2,      &h->hist,      &h->sort_status,   &h->scalars,     &h->pos_sorted, &h->out_key,
                     &h->out_idx,    &h->node_start, &h->node_start_next, &h->selbits, &h->tile_sel, &h->child_count, &h->tile_rank0,   &h->sel,         &h->scan_status, &h->node_index,
                     &h->node_first, &h->bins,      &h->ids_tmp,     &h->dense_counts, &h->node_gcount,
                     &h->part_tile_counts, &h->part_send_counts };
  for (DevBuf* b : bufs)
    b->release();
  release_store(h);
  for (int i = 0; i < 4; ++i) {
    h->list_key[i].release();
    h->list_idx[i].release();
  }
  DevBuf* face_bufs[] = { &h->face_flags, &h->face_offs, &h->face_scan, &h->face_rec, &h->face_src };
  for (DevBuf* b : face_bufs)
    b->release();
  DevBuf* st_bufs[] = { &h->st_slot, &h->st_cnt,  &h->st_boff,   &h->st_scan, &h->st_gcount, &h->st_lo,    &h->st_found,
                        &h->st_cumf, &h->st_nidx, &h->st_ncnt,   &h->st_nflags, &h->st_nsrc, &h->st_nfirst, &h->st_nids };
  for (DevBuf* b : st_bufs)
    b->release();
  free_min_distance_scratch(h->md);
  if (h->h_scalars)
    cudaFreeHost(h->h_scalars);
  for (auto& e : h->ev)
    if (e)
      cudaEventDestroy(e);
  delete h;
}

const char*
swgpu_last_error(swgpu_handle h)
{
  return h ? h->err.c_str() : "invalid handle";
}

int
swgpu_set_stream(swgpu_handle h, void* cuda_stream)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  return SW_OK;
}

int
swgpu_set_deep_node_policy(swgpu_handle h, int policy)
{
  if (!h || policy < 0 || policy > 1)
    return SW_ERR_INVALID_ARGUMENT;
  h->deep_policy = policy;
  return SW_OK;
}

int
swgpu_set_multi_batch(swgpu_handle h, int enable)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  release_store(h);
  h->multi_batch = enable != 0;
  h->batch_done = false;
  h->finalized = false;
  h->out_count = 0;
  h->node_count = 0;
  h->chunks.clear();
  return SW_OK;
}

int
swgpu_reserve(swgpu_handle h, uint64_t n)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  return ensure_batch_buffers(h, n);
}

// the indexing kernels move 16-byte vectors (two points = three double2, four LAS records = three int4)
static bool
aligned16(const void* p)
{
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

int
swgpu_index_batch_device(swgpu_handle h, double* xyz_device, uint64_t n)
{
  if (!h || (!xyz_device && n))
    return SW_ERR_INVALID_ARGUMENT;
  if (!aligned16(xyz_device))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "device positions must be 16-byte aligned");
  cudaSetDevice(h->device);
  h->d_las = nullptr;
  h->d_xyz = xyz_device;
  h->n = n;
  return tile_current_batch(h);
}

int
swgpu_index_batch(swgpu_handle h, double* xyz_host, uint64_t n)
{
  if (!h || (!xyz_host && n))
    return SW_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  CK(h->xyz_own.ensure(std::max<size_t>(n, 1) * 24));
  CK(cudaMemcpyAsync(h->xyz_own.p, xyz_host, n * 24, cudaMemcpyHostToDevice, h->stream));
  h->d_las = nullptr;
  h->d_xyz = h->xyz_own.as<double>();
  h->n = n;
  const int rc = tile_current_batch(h);
  if (rc)
    return rc;
  if (h->n_clamped) { // index_point wrote clamped coordinates back into the PointBuffer
    CK(cudaMemcpyAsync(xyz_host, h->xyz_own.p, n * 24, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return SW_OK;
}

int
swgpu_finalize(swgpu_handle h)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  return h->multi_batch ? run_finalize_store(h) : run_finalize(h);
}

static int
set_las_transform(swgpu_tiler* h, const sw_las_transform* t)
{
  for (int a = 0; a < 3; ++a) {
    h->las_t.scale[a] = t->scale[a];
    h->las_t.offset[a] = t->offset[a];
    h->las_t.hmin[a] = t->header_min[a];
    h->las_t.hmax[a] = t->header_max[a];
    h->las_t.center[a] = t->center[a];
  }
  h->las_t.shift = t->shift_to_center != 0;
  return SW_OK;
}

int
swgpu_index_batch_las_device(swgpu_handle h, const int32_t* las_xyz_device, uint64_t n, const sw_las_transform* t)
{
  if (!h || !t || (!las_xyz_device && n))
    return SW_ERR_INVALID_ARGUMENT;
  if (!aligned16(las_xyz_device))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "device LAS records must be 16-byte aligned");
  cudaSetDevice(h->device);
  set_las_transform(h, t);
  CK(h->xyz_own.ensure(std::max<size_t>(n, 1) * 24));
  h->d_las = las_xyz_device;
  h->d_xyz = h->xyz_own.as<double>();
  h->n = n;
  const int rc = tile_current_batch(h);
  h->d_las = nullptr; // the positions now live in xyz_own
  return rc;
}

int
swgpu_index_batch_las(swgpu_handle h, const int32_t* las_xyz_host, uint64_t n, const sw_las_transform* t)
{
  if (!h || !t || (!las_xyz_host && n))
    return SW_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  CK(h->las_own.ensure(std::max<size_t>(n, 1) * 12));
  CK(cudaMemcpyAsync(h->las_own.p, las_xyz_host, n * 12, cudaMemcpyHostToDevice, h->stream));
  return swgpu_index_batch_las_device(h, h->las_own.as<int32_t>(), n, t);
}

int
swgpu_get_positions(swgpu_handle h, double* xyz_host)
{
  if (!h || !xyz_host)
    return SW_ERR_INVALID_ARGUMENT;
  if (!h->batch_done)
    return fail(h, SW_ERR_STATE, "no batch has been indexed");
  cudaSetDevice(h->device);
  CK(cudaMemcpyAsync(xyz_host, h->d_xyz, h->n * 24, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SW_OK;
}

int
swgpu_result_size(swgpu_handle h, uint64_t* n_nodes, uint64_t* n_point_ids)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  if (n_nodes)
    *n_nodes = h->multi_batch ? store_node_total(h) : h->node_count;
  if (n_point_ids)
    *n_point_ids = h->multi_batch ? store_id_total(h) : h->out_count;
  return SW_OK;
}

// node-major ids: sorted position -> original index (-> global id when this handle tiles a shard)
static void
compose_output_ids(swgpu_tiler* h, u32* out_device)
{
  if (h->out_idx_original) { // the lists carried original indices all along
    if (h->global_ids)
      launch_compose_ids(h->global_ids, h->out_idx.as<u32>(), h->out_count, out_device, h->stream);
    else
      cudaMemcpyAsync(out_device, h->out_idx.p, h->out_count * 4, cudaMemcpyDeviceToDevice, h->stream);
    return;
  }
  if (h->global_ids)
    launch_compose_ids_mapped(h->vals[0].as<u32>(), h->out_idx.as<u32>(), h->global_ids, h->out_count, out_device,
                              h->stream);
  else
    launch_compose_ids(h->vals[0].as<u32>(), h->out_idx.as<u32>(), h->out_count, out_device, h->stream);
}

static int
fill_node_table(swgpu_tiler* h, sw_node* nodes)
{
  if (!nodes || !h->node_count)
    return SW_OK;
  std::vector<u64> index(h->node_count), first(h->node_count);
  CK(cudaMemcpyAsync(index.data(), h->node_index.p, h->node_count * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(first.data(), h->node_first.p, h->node_count * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (const Chunk& c : h->chunks) {
    for (u32 k = 0; k < c.n_nodes; ++k) {
      sw_node& nd = nodes[c.node_base + k];
      nd.index = index[c.node_base + k];
      nd.levels = (uint32_t)c.levels;
      const u64 fmask = ~(1ull << 63); // bit 63 of node_first = take-all marker of the compaction kernel
      nd.flags = c.flags | ((first[c.node_base + k] >> 63) ? SW_NODE_TAKE_ALL : 0u);
      nd.first = first[c.node_base + k] & fmask;
      const u64 end = (k + 1 < c.n_nodes) ? (first[c.node_base + k + 1] & fmask) : (c.out_offset + c.count);
      nd.count = end - nd.first;
    }
  }
  return SW_OK;
}

int
swgpu_get_nodes_device_ids(swgpu_handle h, sw_node* nodes, uint32_t* point_ids_device)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  if (!h->batch_done)
    return fail(h, SW_ERR_STATE, "no batch has been indexed");
  cudaSetDevice(h->device);
  if (h->multi_batch)
    return store_get_nodes(h, nodes, point_ids_device, true);
  if (point_ids_device && h->out_count) {
    compose_output_ids(h, point_ids_device);
    CK(cudaGetLastError());
  }
  return fill_node_table(h, nodes);
}

int
swgpu_get_nodes(swgpu_handle h, sw_node* nodes, uint32_t* point_ids)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  if (!h->batch_done)
    return fail(h, SW_ERR_STATE, "no batch has been indexed");
  cudaSetDevice(h->device);
  if (h->multi_batch)
    return store_get_nodes(h, nodes, point_ids, false);
  if (point_ids && h->out_count) {
    CK(h->ids_tmp.ensure(h->out_count * 4));
    compose_output_ids(h, h->ids_tmp.as<u32>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(point_ids, h->ids_tmp.p, h->out_count * 4, cudaMemcpyDeviceToHost, h->stream));
  }
  const int rc = fill_node_table(h, nodes);
  if (rc)
    return rc;
  CK(cudaStreamSynchronize(h->stream));
  return SW_OK;
}

// get_bounds_from_node_index: get_octant_bounds iterated from the root (tiling/OctreeAlgorithms.cpp:3-18,
// 64-72); the recurrence, not a closed form, so the doubles match the reference bit for bit
static void
node_bounds_of(const swgpu_tiler* h, u64 index, u32 levels, double mn[3], double mx[3])
{
  for (int a = 0; a < 3; ++a) {
    mn[a] = h->prm.bounds_min[a];
    mx[a] = h->prm.bounds_max[a];
  }
  for (u32 l = 0; l < levels; ++l) {
    const u32 octant = (u32)((index >> (3 * (levels - 1 - l))) & 7);
    const u32 bit[3] = { (octant >> 2) & 1u, (octant >> 1) & 1u, octant & 1u };
    for (int a = 0; a < 3; ++a) {
      const double half = (mx[a] - mn[a]) / 2;
      if (bit[a])
        mn[a] = mn[a] + half;
      mx[a] = mn[a] + half;
    }
  }
}

// compute_las_scale_from_bounds, io/LASPersistence.cpp:17-28
static double
las_scale_from_bounds(const double mn[3], const double mx[3])
{
  const double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
  const double diagonal = std::sqrt(ex * ex + ey * ey + ez * ez); // Vector3::length, math/Vector3.h
  if (diagonal > 1'000'000)
    return 0.01;
  else if (diagonal > 100'000)
    return 0.001;
  else if (diagonal > 1)
    return 0.001;
  return 0.0001;
}

// The output holds sorted positions.  Strategies that read positions have them gathered into Morton
// order already (K4): node contents are then nearly sequential in that copy.  RANDOM_GRID never gathers, so
// its payload goes through the sort permutation into the caller's array.
static void
payload_source(const swgpu_tiler* h, const double** xyz, const u32** perm)
{
  if (needs_positions(h->prm.sampling) && h->pos_sorted.p) {
    *xyz = h->pos_sorted.as<double>();
    *perm = nullptr;
  } else {
    *xyz = h->d_xyz;
    *perm = h->out_idx_original ? nullptr : h->vals[0].as<u32>(); // the lists may hold original indices already
  }
}

static int
payload_checks(swgpu_tiler* h)
{
  if (!h->batch_done)
    return fail(h, SW_ERR_STATE, "no batch has been indexed");
  if (!h->multi_batch && !h->d_xyz && h->n)
    return fail(h, SW_ERR_STATE, "the positions of the batch are gone");
  return SW_OK;
}

int
swgpu_get_payload_pnts_device(swgpu_handle h, float* xyz_f32_device)
{
  if (!h || !xyz_f32_device)
    return SW_ERR_INVALID_ARGUMENT;
  const int rc = payload_checks(h);
  if (rc)
    return rc;
  cudaSetDevice(h->device);
  if (h->multi_batch) { // node-major global ids of the store, positions of all batches
    const u64 n_ids = store_id_total(h);
    CK(h->ids_tmp.ensure(std::max<u64>(n_ids, 1) * 4));
    const int rc2 = store_get_nodes(h, nullptr, h->ids_tmp.as<u32>(), true);
    if (rc2)
      return rc2;
    launch_payload_pnts(h->store_xyz.as<double>(), nullptr, h->ids_tmp.as<u32>(), n_ids, xyz_f32_device, h->stream);
    CK(cudaGetLastError());
    return SW_OK;
  }
  const double* src = nullptr;
  const u32* perm = nullptr;
  payload_source(h, &src, &perm);
  launch_payload_pnts(src, perm, h->out_idx.as<u32>(), h->out_count, xyz_f32_device, h->stream);
  CK(cudaGetLastError());
  return SW_OK;
}

int
swgpu_get_payload_pnts(swgpu_handle h, float* xyz_f32_host)
{
  if (!h || !xyz_f32_host)
    return SW_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  const u64 n_out = h->multi_batch ? store_id_total(h) : h->out_count;
  CK(h->payload_tmp.ensure(std::max<u64>(n_out, 1) * 12));
  const int rc = swgpu_get_payload_pnts_device(h, h->payload_tmp.as<float>());
  if (rc)
    return rc;
  CK(cudaMemcpyAsync(xyz_f32_host, h->payload_tmp.p, n_out * 12, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SW_OK;
}

int
swgpu_get_payload_las_device(swgpu_handle h, int32_t* xyz_i32_device, sw_las_node_header* headers_host)
{
  if (!h || !xyz_i32_device)
    return SW_ERR_INVALID_ARGUMENT;
  int rc = payload_checks(h);
  if (rc)
    return rc;
  cudaSetDevice(h->device);
  if (h->multi_batch) {
    const u64 n_nodes = store_node_total(h), n_ids = store_id_total(h);
    if (!n_nodes)
      return SW_OK;
    std::vector<sw_node> nodes(n_nodes);
    CK(h->ids_tmp.ensure(std::max<u64>(n_ids, 1) * 4));
    rc = store_get_nodes(h, nodes.data(), h->ids_tmp.as<u32>(), true);
    if (rc)
      return rc;
    std::vector<double> hdr(n_nodes * 4);
    std::vector<u64> first(n_nodes);
    for (u64 row = 0; row < n_nodes; ++row) {
      double mn[3], mx[3];
      node_bounds_of(h, nodes[row].index, nodes[row].levels, mn, mx);
      const double scale = las_scale_from_bounds(mn, mx);
      hdr[4 * row + 0] = mn[0];
      hdr[4 * row + 1] = mn[1];
      hdr[4 * row + 2] = mn[2];
      hdr[4 * row + 3] = scale;
      first[row] = nodes[row].first;
      if (headers_host) {
        sw_las_node_header& o = headers_host[row];
        for (int a = 0; a < 3; ++a) {
          o.offset[a] = mn[a];
          o.max[a] = mx[a];
        }
        o.scale = scale;
        o.reserved = 0;
      }
    }
    CK(h->node_hdr.ensure(n_nodes * 32));
    CK(h->st_nfirst.ensure((n_nodes + 1) * 8));
    CK(cudaMemcpyAsync(h->node_hdr.p, hdr.data(), n_nodes * 32, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->st_nfirst.p, first.data(), n_nodes * 8, cudaMemcpyHostToDevice, h->stream));
    launch_payload_las(h->store_xyz.as<double>(), nullptr, h->ids_tmp.as<u32>(), n_ids, h->st_nfirst.as<u64>(),
                       (u32)n_nodes, h->node_hdr.as<double>(), xyz_i32_device, h->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return SW_OK;
  }
  if (!h->node_count)
    return SW_OK;
  // per-node header values on the host (a few thousand nodes), then one kernel over all points
  std::vector<u64> index(h->node_count);
  CK(cudaMemcpyAsync(index.data(), h->node_index.p, h->node_count * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  std::vector<double> hdr(h->node_count * 4);
  for (const Chunk& c : h->chunks) {
    for (u32 k = 0; k < c.n_nodes; ++k) {
      const u64 row = c.node_base + k;
      double mn[3], mx[3];
      node_bounds_of(h, index[row], (u32)c.levels, mn, mx);
      const double scale = las_scale_from_bounds(mn, mx);
      hdr[4 * row + 0] = mn[0];
      hdr[4 * row + 1] = mn[1];
      hdr[4 * row + 2] = mn[2];
      hdr[4 * row + 3] = scale;
      if (headers_host) {
        sw_las_node_header& o = headers_host[row];
        for (int a = 0; a < 3; ++a) {
          o.offset[a] = mn[a];
          o.max[a] = mx[a];
        }
        o.scale = scale;
        o.reserved = 0;
      }
    }
  }
  CK(h->node_hdr.ensure(h->node_count * 32));
  CK(cudaMemcpyAsync(h->node_hdr.p, hdr.data(), h->node_count * 32, cudaMemcpyHostToDevice, h->stream));
  const double* src = nullptr;
  const u32* perm = nullptr;
  payload_source(h, &src, &perm);
  launch_payload_las(src, perm, h->out_idx.as<u32>(), h->out_count, h->node_first.as<u64>(),
                     (u32)h->node_count, h->node_hdr.as<double>(), xyz_i32_device, h->stream);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream)); // hdr is a stack-owned staging vector
  return SW_OK;
}

int
swgpu_get_payload_las(swgpu_handle h, int32_t* xyz_i32_host, sw_las_node_header* headers_host)
{
  if (!h || !xyz_i32_host)
    return SW_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  const u64 n_out = h->multi_batch ? store_id_total(h) : h->out_count;
  CK(h->payload_tmp.ensure(std::max<u64>(n_out, 1) * 12));
  const int rc = swgpu_get_payload_las_device(h, h->payload_tmp.as<int32_t>(), headers_host);
  if (rc)
    return rc;
  CK(cudaMemcpyAsync(xyz_i32_host, h->payload_tmp.p, n_out * 12, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SW_OK;
}

int
swgpu_get_start_level(swgpu_handle h, int32_t* level)
{
  if (!h || !level)
    return SW_ERR_INVALID_ARGUMENT;
  *level = h->start_level;
  return SW_OK;
}

int
swgpu_get_clamped_count(swgpu_handle h, uint64_t* n)
{
  if (!h || !n)
    return SW_ERR_INVALID_ARGUMENT;
  *n = h->n_clamped;
  return SW_OK;
}

int
swgpu_get_keys(swgpu_handle h, uint64_t* keys, uint32_t* order)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  if (!h->batch_done)
    return fail(h, SW_ERR_STATE, "no batch has been indexed");
  cudaSetDevice(h->device);
  if (keys)
    CK(cudaMemcpyAsync(keys, h->keys[0].p, h->n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (order)
    CK(cudaMemcpyAsync(order, h->vals[0].p, h->n * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SW_OK;
}

int
swgpu_gather_attribute_device(swgpu_handle h, const void* src_device, uint32_t width, void* dst_device)
{
  if (!h || !src_device || !dst_device)
    return SW_ERR_INVALID_ARGUMENT;
  if (!h->batch_done)
    return fail(h, SW_ERR_STATE, "no batch has been indexed");
  cudaSetDevice(h->device);
  const u32* ids = h->out_idx.as<u32>(); // original indices (RANDOM_GRID), else sorted positions to compose
  if (h->multi_batch)
    return fail(h, SW_ERR_STATE, "multi-batch mode: gather by the global ids of swgpu_get_nodes_device_ids instead");
  if (!h->out_idx_original) {
    CK(h->ids_tmp.ensure(std::max<u64>(h->out_count, 1) * 4));
    launch_compose_ids(h->vals[0].as<u32>(), h->out_idx.as<u32>(), h->out_count, h->ids_tmp.as<u32>(), h->stream);
    ids = h->ids_tmp.as<u32>();
  }
  if (width == 24)
    launch_gather_positions(static_cast<const double*>(src_device), ids, h->out_count,
                            static_cast<double*>(dst_device), h->stream);
  else if (width == 1 || width == 2 || width == 3 || width == 4 || width == 8 || width == 12 || width == 16)
    launch_gather_bytes(src_device, ids, h->out_count, width, dst_device, h->stream);
  else
    return fail(h, SW_ERR_INVALID_ARGUMENT, "unsupported attribute width");
  CK(cudaGetLastError());
  return SW_OK;
}

int
swgpu_morton_encode_device(swgpu_handle h, double* xyz_device, uint64_t n, uint64_t* keys_device)
{
  if (!h || (n && (!xyz_device || !keys_device)))
    return SW_ERR_INVALID_ARGUMENT;
  if (!aligned16(xyz_device) || !aligned16(keys_device))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "device positions and keys must be 16-byte aligned");
  cudaSetDevice(h->device);
  CK(h->hist.ensure(sort_hist_words() * 4));
  CK(cudaMemsetAsync(h->hist.p, 0, sort_hist_words() * 4, h->stream));
  CK(cudaMemsetAsync(h->scalars.p, 0, 128, h->stream));
  launch_morton_encode(xyz_device, n, h->bounds, reinterpret_cast<u64*>(keys_device), h->hist.as<u32>(), h->d_n_clamped(), h->stream);
  CK(cudaGetLastError());
  const int rc = sync_scalars(h);
  if (rc)
    return rc;
  h->n_clamped = h->h_scalars->n_clamped;
  return SW_OK;
}

int
swgpu_sort_keys_device(swgpu_handle h, uint64_t* keys_device, uint64_t n, uint32_t* order_device)
{
  if (!h || (n && (!keys_device || !order_device)))
    return SW_ERR_INVALID_ARGUMENT;
  if (n >= (1ull << 30))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "a batch is limited to 2^30 - 1 points per GPU");
  if (n == 0)
    return SW_OK;
  cudaSetDevice(h->device);
  CK(h->keys[1].ensure(n * 8));
  CK(h->vals[1].ensure(n * 4));
  CK(h->hist.ensure(sort_hist_words() * 4));
  CK(h->sort_status.ensure(sort_status_words(n) * 4));
  CK(cudaMemsetAsync(h->hist.p, 0, sort_hist_words() * 4, h->stream));
  launch_key_histogram(reinterpret_cast<const u64*>(keys_device), n, h->hist.as<u32>(), h->stream);
  // an explicit swgpu_set_sort_mode applies here too (keys with bit 63 clear); automatic mode sorts all digits,
  // arbitrary keys carry no density hint
  const int fp = h->sort_mode > 0 ? h->sort_mode : 0;
  if (sort_input_buffer_top(fp) == 1) // odd number of passes: start in the scratch buffer, finish in the caller's
    CK(cudaMemcpyAsync(h->keys[1].p, keys_device, n * 8, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemsetAsync(h->d_sort_stats(), 0, 24, h->stream));
  const int rc = sort_pairs(h, reinterpret_cast<u64*>(keys_device), h->keys[1].as<u64>(), order_device,
                            h->vals[1].as<u32>(), n, fp, false);
  if (rc)
    return rc;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return SW_OK;
}

int
swgpu_prefix_histogram_device(swgpu_handle h, const uint64_t* keys_device, uint64_t n, uint32_t* bins_device)
{
  if (!h || !bins_device || (n && !keys_device))
    return SW_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  launch_prefix_histogram(reinterpret_cast<const u64*>(keys_device), n, bins_device, h->stream);
  CK(cudaGetLastError());
  return SW_OK;
}

int
swgpu_prefix_histogram_coarse_device(swgpu_handle h, const uint64_t* keys_device, uint64_t n, uint32_t levels,
                                     uint32_t* bins_device)
{
  if (!h || !bins_device || (n && !keys_device) || levels < 1 || levels > 4)
    return SW_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  launch_prefix_histogram_coarse(reinterpret_cast<const u64*>(keys_device), n, (int)levels, bins_device, h->stream);
  CK(cudaGetLastError());
  return SW_OK;
}

int
swgpu_estimate_start_level(const uint32_t* bins_host, uint32_t concurrency, int32_t* level)
{
  if (!bins_host || !level)
    return SW_ERR_INVALID_ARGUMENT;
  std::vector<u64> prefix(262145, 0);
  for (u32 b = 0; b < 262144u; ++b)
    prefix[b + 1] = prefix[b] + bins_host[b];
  *level = start_level_from_prefix_counts([&](u32 b, u32 group) -> u64 { return prefix[b + group] - prefix[b]; },
                                          concurrency ? concurrency : 1u);
  return SW_OK;
}

int
swgpu_choose_splitters(const uint32_t* bins_host, uint32_t n_ranks, uint32_t shard_levels, uint32_t* first_prefix)
{
  if (!bins_host || !first_prefix || n_ranks == 0 || n_ranks > SWGPU_MAX_RANKS || shard_levels == 0 || shard_levels > 6)
    return SW_ERR_INVALID_ARGUMENT;
  const u32 group = 1u << (3 * (6 - shard_levels)); // level-5 prefixes per shard-level subtree
  const u32 n_groups = 262144u / group;
  std::vector<u64> cum(n_groups + 1, 0);
  for (u32 g = 0; g < n_groups; ++g) {
    u64 c = 0;
    for (u32 b = 0; b < group; ++b)
      c += bins_host[g * group + b];
    cum[g + 1] = cum[g] + c;
  }
  const u64 total = cum[n_groups];
  first_prefix[0] = 0;
  u32 g = 0;
  for (u32 r = 1; r < n_ranks; ++r) {
    // boundary whose cumulative count is closest to r/n_ranks of the points, never moving backwards
    const double target = (double)total * r / n_ranks;
    while (g < n_groups && (double)cum[g + 1] <= target)
      ++g;
    u32 pick = g;
    if (g < n_groups && (double)cum[g + 1] - target < target - (double)cum[g])
      pick = g + 1;
    if (pick * group < first_prefix[r - 1])
      pick = first_prefix[r - 1] / group;
    first_prefix[r] = pick * group;
  }
  first_prefix[n_ranks] = 262144u;
  return SW_OK;
}

int
swgpu_max_shard_levels(swgpu_handle h, uint32_t* shard_levels)
{
  if (!h || !shard_levels)
    return SW_ERR_INVALID_ARGUMENT;
  int depth = 6;
  switch (h->prm.sampling) {
    case SW_RANDOM_GRID:
    case SW_GRID_CENTER:
      // root node cells are prefixes of cand(-1) + 1 levels; deeper nodes use deeper cells
      depth = std::min(6, cand_level_sampler(h, -1) + 1);
      break;
    case SW_JITTERED: { // root grid: log2(cells) levels (Sampling.h:621-660)
      const double perfect = root_extent_x(h) / (double)h->prm.spacing_at_root;
      const uint32_t cells = prev_pow2_u32(static_cast<uint32_t>(perfect));
      depth = std::min(6, cells ? (int)std::log2(cells) : 0);
      break;
    }
    default: { // MIN_DISTANCE: no cell structure; nodes above the shard depth are sampled per shard and the
               // shard faces resolved afterwards (swgpu_set_shard_faces), which needs shard subtrees that are at
               // least one spacing wide
      const double ratio = root_extent_x(h) / ((double)h->prm.spacing_at_root * (1.0 + 1e-6));
      depth = std::min(6, ratio >= 1.0 ? (int)std::floor(std::log2(ratio)) : 0);
      break;
    }
  }
  *shard_levels = (uint32_t)std::max(depth, 0);
  return SW_OK;
}

int
swgpu_partition_device(swgpu_handle h, const uint64_t* keys_device, const double* xyz_device, uint64_t n,
                       const uint32_t* first_prefix, uint32_t n_ranks, uint32_t id_base, double* out_xyz_device,
                       uint32_t* out_id_device, uint64_t* send_counts_host)
{
  if (!h || !first_prefix || !send_counts_host || n_ranks == 0 || n_ranks > SWGPU_MAX_RANKS ||
      (n && (!keys_device || !xyz_device || !out_xyz_device || !out_id_device)))
    return SW_ERR_INVALID_ARGUMENT;
  if (n >= (1ull << 32) || (u64)id_base + n > (1ull << 32))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "global point ids are 32 bit: id_base + n must not exceed 2^32");
  cudaSetDevice(h->device);
  CK(h->part_tile_counts.ensure(partition_tiles(n) * SW_MAX_RANKS * 4));
  CK(h->part_send_counts.ensure(SW_MAX_RANKS * 8));
  launch_partition_by_splitters(reinterpret_cast<const u64*>(keys_device), xyz_device, n, first_prefix, n_ranks, id_base,
                                h->part_tile_counts.as<u32>(), h->part_send_counts.as<u64>(), out_xyz_device,
                                out_id_device, static_cast<const u32*>(h->part_attr_src), h->part_attr_words,
                                static_cast<u32*>(h->part_attr_dst), h->stream);
  CK(cudaGetLastError());
  u64 counts[SW_MAX_RANKS];
  CK(cudaMemcpyAsync(counts, h->part_send_counts.p, SW_MAX_RANKS * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (u32 r = 0; r < n_ranks; ++r)
    send_counts_host[r] = counts[r];
  return SW_OK;
}

int
swgpu_partition_to_peers_device(swgpu_handle h, const uint64_t* keys_device, const double* xyz_device, uint64_t n,
                                const uint32_t* first_prefix, uint32_t n_ranks, uint32_t id_base,
                                void* const* peer_xyz_device, void* const* peer_ids_device, const uint64_t* dst_offsets,
                                uint64_t* send_counts_host)
{
  if (!h || !first_prefix || !peer_xyz_device || !peer_ids_device || !dst_offsets || n_ranks == 0 ||
      n_ranks > SWGPU_MAX_RANKS || (n && (!keys_device || !xyz_device)))
    return SW_ERR_INVALID_ARGUMENT;
  if (n >= (1ull << 32) || (u64)id_base + n > (1ull << 32))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "global point ids are 32 bit: id_base + n must not exceed 2^32");
  cudaSetDevice(h->device);
  CK(h->part_tile_counts.ensure(partition_tiles(n) * SW_MAX_RANKS * 4));
  CK(h->part_send_counts.ensure(SW_MAX_RANKS * 8));
  double* px[SW_MAX_RANKS];
  u32* pi[SW_MAX_RANKS];
  u64 off[SW_MAX_RANKS];
  for (u32 r = 0; r < n_ranks; ++r) {
    px[r] = static_cast<double*>(peer_xyz_device[r]);
    pi[r] = static_cast<u32*>(peer_ids_device[r]);
    off[r] = dst_offsets[r];
  }
  u32* pa[SW_MAX_RANKS];
  for (u32 r = 0; r < n_ranks; ++r)
    pa[r] = h->part_attr_peers ? static_cast<u32*>(h->part_attr_peers[r]) : nullptr;
  launch_partition_to_peers(reinterpret_cast<const u64*>(keys_device), xyz_device, n, first_prefix, n_ranks, id_base,
                            h->part_tile_counts.as<u32>(), h->part_send_counts.as<u64>(), px, pi, off,
                            static_cast<const u32*>(h->part_attr_src), h->part_attr_words,
                            h->part_attr_peers ? pa : nullptr, h->stream);
  CK(cudaGetLastError());
  if (send_counts_host) { // optional check value; costs a stream synchronisation
    u64 counts[SW_MAX_RANKS];
    CK(cudaMemcpyAsync(counts, h->part_send_counts.p, SW_MAX_RANKS * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (u32 r = 0; r < n_ranks; ++r)
      send_counts_host[r] = counts[r];
  }
  return SW_OK;
}

int
swgpu_set_shard(swgpu_handle h, uint32_t shard_levels, int32_t start_level, swgpu_allreduce_u32_fn allreduce,
                void* allreduce_ctx, const uint32_t* global_ids_device)
{
  if (!h || shard_levels > 6)
    return SW_ERR_INVALID_ARGUMENT;
  if (shard_levels && h->prm.max_points_per_node >= (1ull << 28))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "sharded tiling needs max_points_per_node < 2^28");
  h->shard_levels = shard_levels;
  h->start_level_override = shard_levels ? start_level : -1;
  h->allreduce = shard_levels ? allreduce : nullptr;
  h->allreduce_ctx = allreduce_ctx;
  h->global_ids = shard_levels ? reinterpret_cast<const u32*>(global_ids_device) : nullptr;
  return SW_OK;
}

int
swgpu_set_partition_attributes(swgpu_handle h, const void* attr_device, uint32_t attr_bytes, void* out_attr_device,
                               void* const* peer_attr_device)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  if (!attr_device) {
    h->part_attr_src = nullptr;
    h->part_attr_words = 0;
    h->part_attr_dst = nullptr;
    h->part_attr_peers = nullptr;
    return SW_OK;
  }
  if (attr_bytes == 0 || attr_bytes > 16 || (attr_bytes & 3u))
    return fail(h, SW_ERR_INVALID_ARGUMENT, "attribute records must be 4, 8, 12 or 16 bytes per point");
  h->part_attr_src = attr_device;
  h->part_attr_words = attr_bytes / 4;
  h->part_attr_dst = out_attr_device;
  h->part_attr_peers = peer_attr_device;
  return SW_OK;
}

int
swgpu_set_shard_faces(swgpu_handle h, const uint32_t* first_prefix, uint32_t n_ranks, uint32_t rank,
                      swgpu_allgatherv_fn allgatherv, void* ctx)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  if (!allgatherv) {
    h->face_fn = nullptr;
    h->face_ctx = nullptr;
    return SW_OK;
  }
  if (!first_prefix || n_ranks == 0 || n_ranks > SWGPU_MAX_RANKS || rank >= n_ranks)
    return SW_ERR_INVALID_ARGUMENT;
  for (u32 r = 0; r <= n_ranks; ++r)
    h->face_first_prefix[r] = first_prefix[r];
  h->face_n_ranks = n_ranks;
  h->face_rank = rank;
  h->face_fn = allgatherv;
  h->face_ctx = ctx;
  return SW_OK;
}

int
swgpu_set_sort_mode(swgpu_handle h, int mode)
{
  if (!h || mode < -1 || mode > 3)
    return SW_ERR_INVALID_ARGUMENT;
  h->sort_mode = mode;
  h->sort_next = 0;
  h->sort_probe_wait = 0;
  return SW_OK;
}

int
swgpu_enable_timing(swgpu_handle h, int enable)
{
  if (!h)
    return SW_ERR_INVALID_ARGUMENT;
  h->timing = enable != 0;
  return SW_OK;
}

int
swgpu_get_stats(swgpu_handle h, swgpu_stats* out)
{
  if (!h || !out)
    return SW_ERR_INVALID_ARGUMENT;
  h->stats.n_output_ids = h->out_count;
  h->stats.n_nodes = h->node_count;
  if (h->timing && h->batch_done) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaEventElapsedTime(&h->stats.ms_index, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->stats.ms_sort, h->ev[1], h->ev[2]);
    h->stats.ms_sort_finish = 0.f;
    if (h->stats.sort_first_bit && h->stats.n_points && h->stats.sort_fallback != 1) {
      float before = 0.f;
      cudaEventElapsedTime(&before, h->ev[1], h->ev[6]);
      h->stats.ms_sort_finish = h->stats.ms_sort - before;
    }
    cudaEventElapsedTime(&h->stats.ms_gather, h->ev[2], h->ev[3]);
    cudaEventElapsedTime(&h->stats.ms_sample, h->ev[3], h->ev[4]);
    const bool fin = h->finalized && h->prm.tiling == SW_FAST;
    cudaEventElapsedTime(&h->stats.ms_total, h->ev[0], fin ? h->ev[5] : h->ev[4]);
    if (fin) {
      float r = 0.f;
      cudaEventElapsedTime(&r, h->ev[4], h->ev[5]);
      h->stats.ms_sample += r;
    }
    cudaGetLastError();
  }
  *out = h->stats;
  return SW_OK;
}

} // extern "C"
