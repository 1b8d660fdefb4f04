// TilingAlgorithmGPU.h — drop-in TilingAlgorithmBase for the reference tree.
//
// Copy this file and swgpu_tiler.hpp into schwarzwald/core/tiling/, link libswgpu.so, and pick it
// in Tiler::Tiler next to TilingAlgorithmV1 / V3 (core/process/Tiler.cpp:189-198); INTEGRATION.md
// shows the three-line patch.  It compiles against the reference's own headers (taskflow, Boost,
// GSL) — none of which exist in this repository's build image, so the repository checks it against
// tests/host_mock/ (signature-compatible stand-ins) with `g++ -fsyntax-only`.
//
// What it replaces, call for call (paths relative to schwarzwald/core):
//   TilingAlgorithmV1::build_execution_graph   tiling/TilingAlgorithms.cpp:577-626   (ACCURATE)
//   TilingAlgorithmV3::build_execution_graph   tiling/TilingAlgorithms.cpp:1250-1360 (FAST)
//   TilingAlgorithmV3::finalize                tiling/TilingAlgorithms.cpp:1239-1248, 1717-1784
//   tile_terminal_node / tile_internal_node    the persist_points + increment_progress calls at
//                                              :232-240 and :330-345
//   tile_node with cached points               :351-492 / read_pnts_from_disk :50-109 (several batches)
// Batches: Tiler::run hands over at most internal_cache_size points per call.  A first batch that is shorter
// than that is the whole cloud: it runs through the single-batch pipeline and is handed to the sink at once.
// A full first batch may be followed by more: the library then keeps every node's points in device memory
// (swgpu_set_multi_batch), later batches are merged with them exactly like tile_node does with the points it
// reads back from the persistence, and the FINAL content of every node is handed to the sink once, in
// finalize() — the reference rewrites a node's file on every visit (PointsPersistence.h:23-31 "overwrites"),
// so the files end up the same.  The adapter keeps a host copy of every batch (all attributes) for that
// hand-off, because Tiler reuses its two PointBuffers (process/Tiler.h:135).
//
// Lossy sinks (LAS / LAZ / ENTWINE_*: PointsPersistence::is_lossless() == false).  Where the reference READS
// POINTS BACK from the persistence it sees quantised, clamped positions and re-sorts them: reconstruct_single_node
// (FAST, TilingAlgorithms.cpp:1661-1691) and read_pnts_from_disk (several batches, :50-109).  This adapter samples
// those nodes from the original doubles still held in device memory, so for FAST or multi-batch runs into a
// lossy sink the reconstructed / revisited nodes can differ from the reference's in points whose quantised
// position changes a cell winner (the nodes at and below the start level of a single-batch run are identical).
// The adapter says so once on stderr, or throws when constructed with refuse_lossy_read_back = true.
// Lossless sinks (3DTILES, BINARY, in-memory) are bit-identical in every mode.
#pragma once

#include "swgpu_tiler.hpp"

#include "datastructures/PointBuffer.h"
#include "io/PointsPersistence.h"
#include "math/AABB.h"
#include "tiling/Sampling.h"
#include "tiling/TilingAlgorithms.h"

#include <cstdio>
#include <memory>
#include <optional>
#include <string>
#include <variant>
#include <vector>

struct TilingAlgorithmGPU : TilingAlgorithmBase
{
  TilingAlgorithmGPU(SamplingStrategy& sampling_strategy,
                     ProgressReporter* progress_reporter,
                     PointsPersistence& persistence,
                     TilerMetaParameters meta_parameters,
                     int cuda_device = 0,
                     bool refuse_lossy_read_back = false)
    : TilingAlgorithmBase(sampling_strategy, progress_reporter, persistence, meta_parameters)
    , _cuda_device(cuda_device)
    , _refuse_lossy_read_back(refuse_lossy_read_back)
  {}

  // All GPUs of the box from this one process: a batch that is the whole cloud (shorter than
  // internal_cache_size) is sharded over `cuda_devices` (swgpu_multi_*); a multi-batch run stays on the first one,
  // because the node store between batches is per GPU.
  TilingAlgorithmGPU(SamplingStrategy& sampling_strategy,
                     ProgressReporter* progress_reporter,
                     PointsPersistence& persistence,
                     TilerMetaParameters meta_parameters,
                     std::vector<int> cuda_devices,
                     bool refuse_lossy_read_back = false)
    : TilingAlgorithmBase(sampling_strategy, progress_reporter, persistence, meta_parameters)
    , _cuda_device(cuda_devices.empty() ? 0 : cuda_devices.front())
    , _refuse_lossy_read_back(refuse_lossy_read_back)
    , _cuda_devices(std::move(cuda_devices))
  {}

  std::pair<tf::Task, tf::Task> build_execution_graph(util::Range<PointBuffer::PointIterator> points,
                                                      const AABB& bounds,
                                                      uint32_t num_indexing_threads,
                                                      tf::Taskflow& tf) override
  {
    // one task, like the single-threaded sort task of V1 (TilingAlgorithms.cpp:600-604); the GPU does
    // index + sort + every sampling level inside it
    auto task = tf.emplace([this, points, bounds, num_indexing_threads]() mutable {
      const auto n = static_cast<uint64_t>(points.size());
      if (!_batches_done++) {
        _root_bounds = bounds;
        const double bmin[3] = { bounds.min.x, bounds.min.y, bounds.min.z };
        const double bmax[3] = { bounds.max.x, bounds.max.y, bounds.max.z };
        if (_cuda_devices.size() > 1 && n < _meta_parameters.internal_cache_size) {
          // the whole cloud in one batch: every GPU tiles its Morton-prefix shard
          swgpu::MultiTiler multi(sampling_enum(),
                                  _meta_parameters.tiling_strategy == TilingStrategy::Fast ? SW_FAST : SW_ACCURATE,
                                  _meta_parameters.spacing_at_root,
                                  _meta_parameters.max_depth,
                                  _meta_parameters.max_points_per_node,
                                  bmin,
                                  bmax,
                                  num_indexing_threads,
                                  _cuda_devices);
          warn_if_lossy(false);
          double* xyz_all = n ? &(*std::begin(points)).position().x : nullptr;
          multi.index_batch(xyz_all, n);
          multi.finalize();
          hand_off(multi.result(), std::begin(points), /*count_progress=*/true);
          _done_on_many_gpus = true;
          return;
        }
        _tiler = std::make_unique<swgpu::Tiler>(sampling_enum(),
                                                _meta_parameters.tiling_strategy == TilingStrategy::Fast ? SW_FAST
                                                                                                         : SW_ACCURATE,
                                                _meta_parameters.spacing_at_root,
                                                _meta_parameters.max_depth,
                                                _meta_parameters.max_points_per_node,
                                                bmin,
                                                bmax,
                                                num_indexing_threads,
                                                _cuda_device);
        // nodes where the reference would re-root (TilingAlgorithms.cpp:444-483) are stored whole instead of failing
        // the run; hand_off() reports them
        _tiler->set_deep_node_policy(true);
        // a full batch may be followed by more (Tiler::run fills internal_cache_size points per batch)
        _multi_batch = n >= _meta_parameters.internal_cache_size;
        if (_multi_batch)
          _tiler->set_multi_batch(true);
        warn_if_lossy(_multi_batch);
      } else if (_done_on_many_gpus || !_multi_batch) {
        throw std::runtime_error{ "TilingAlgorithmGPU: a batch followed a batch shorter than internal_cache_size" };
      }
      // PointBuffer::positions() is a std::vector<Vector3<double>>: AoS x,y,z doubles (PointBuffer.h:291)
      double* xyz = n ? &(*std::begin(points)).position().x : nullptr;
      _tiler->index_batch(xyz, n); // clamps outliers in place, as index_point does
      if (_multi_batch) {
        // keep the batch (positions after clamping, every attribute) for the hand-off in finalize()
        std::vector<PointBuffer::PointReference> refs(std::begin(points), std::end(points));
        _stored_points.append_buffer(PointBuffer{ gsl::span<PointBuffer::PointReference>{ refs.data(), refs.size() } });
        // every point of the batch ends up in exactly one node (TilingAlgorithms.cpp:238-240, 336-345)
        if (_progress_reporter)
          _progress_reporter->increment_progress<size_t>(progress::INDEXING, n);
        return;
      }
      // a batch shorter than internal_cache_size is the only one: FAST's reconstruction of the upper levels
      // (TilingAlgorithmV3::finalize) runs right here, while the caller's PointBuffer is still alive
      _tiler->finalize();
      hand_off(_tiler->result(), std::begin(points), /*count_progress=*/true);
    });
    return { task, task };
  }

  void finalize(const AABB& bounds) override
  {
    if (!_tiler)
      return;
    if (_multi_batch) {
      _tiler->finalize();
      hand_off(_tiler->result(), std::begin(_stored_points), /*count_progress=*/false);
      return;
    }
  }

private:
  sw_sampling sampling_enum() const
  {
    return std::visit(
      [](const auto& s) -> sw_sampling {
        using T = std::decay_t<decltype(s)>;
        if constexpr (std::is_same_v<T, RandomSortedGridSampling>)
          return SW_RANDOM_GRID;
        else if constexpr (std::is_same_v<T, GridCenterSampling>)
          return SW_GRID_CENTER;
        else if constexpr (std::is_same_v<T, PoissonDiskSampling>)
          return SW_MIN_DISTANCE;
        else if constexpr (std::is_same_v<T, JitteredSampling>)
          return SW_JITTERED;
        else // AdaptivePoissonDiskSampling: its density function is an opaque std::function; the only
             // one the reference ever constructs is the CLI's (process/TilerProcess.cpp:500-508),
             // which is what SW_MIN_DISTANCE_FAST implements
          return SW_MIN_DISTANCE_FAST;
      },
      _sampling_strategy);
  }

  // The per-node persist_points calls of tile_terminal_node / tile_internal_node /
  // reconstruct_single_node, driven from the node table instead of the recursion.
  void warn_if_lossy(bool multi_batch)
  {
    if (_persistence.is_lossless() || !(multi_batch || _meta_parameters.tiling_strategy == TilingStrategy::Fast))
      return;
    const char* msg = "TilingAlgorithmGPU: the output format is lossy; nodes the reference would re-read from "
                      "disk (FAST reconstruction, later batches) are sampled from the original positions "
                      "instead of the quantised ones";
    if (_refuse_lossy_read_back)
      throw std::runtime_error{ msg };
    std::fprintf(stderr, "warning: %s\n", msg);
  }

  void hand_off(const swgpu::Tiler::Result& result, PointBuffer::PointIterator first_point, bool count_progress)
  {
    const double rmin[3] = { _root_bounds.min.x, _root_bounds.min.y, _root_bounds.min.z };
    const double rmax[3] = { _root_bounds.max.x, _root_bounds.max.y, _root_bounds.max.z };
    std::vector<PointBuffer::PointReference> refs;
    size_t deep_nodes = 0, deep_points = 0;
    for (const sw_node& node : result.nodes) {
      if (node.flags & SW_NODE_DEEP) {
        ++deep_nodes;
        deep_points += node.count;
      }
      refs.clear();
      refs.reserve(node.count);
      for (uint64_t k = 0; k < node.count; ++k)
        refs.push_back(first_point[static_cast<std::ptrdiff_t>(result.point_ids[node.first + k])]);
      double nmin[3], nmax[3];
      swgpu::node_bounds(node.index, node.levels, rmin, rmax, nmin, nmax);
      const AABB node_aabb{ { nmin[0], nmin[1], nmin[2] }, { nmax[0], nmax[1], nmax[2] } };
      _persistence.persist_points(
        std::begin(refs), std::end(refs), node_aabb, swgpu::node_name(node.index, node.levels));
      // progress contract (TilingAlgorithms.cpp:238-240, 336-345): reconstructed copies do not count
      if (count_progress && _progress_reporter && !(node.flags & SW_NODE_RECONSTRUCTED))
        _progress_reporter->increment_progress<size_t>(progress::INDEXING, node.count);
    }
    if (deep_nodes)
      std::fprintf(stderr,
                   "note: %zu nodes at the depth where the reference re-indexes with a new root hold their %zu "
                   "remaining points unsampled (duplicate or extremely dense points)\n",
                   deep_nodes, deep_points);
  }

  int _cuda_device;
  bool _refuse_lossy_read_back;
  size_t _batches_done = 0;
  bool _multi_batch = false;
  bool _done_on_many_gpus = false;
  std::vector<int> _cuda_devices;
  PointBuffer _stored_points; // multi-batch: host copy of every batch, indexed by global point id
  AABB _root_bounds;
  std::unique_ptr<swgpu::Tiler> _tiler;
};
