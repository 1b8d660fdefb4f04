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
// Scope: the single-batch regime (internal_cache_size >= number of points).  A second batch throws,
// because merging with persisted nodes (TilingAlgorithms.cpp:50-109) is not part of this library.
#pragma once

#include "swgpu_tiler.hpp"

#include "datastructures/PointBuffer.h"
#include "io/PointsPersistence.h"
#include "math/AABB.h"
#include "tiling/Sampling.h"
#include "tiling/TilingAlgorithms.h"

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
                     int cuda_device = 0)
    : TilingAlgorithmBase(sampling_strategy, progress_reporter, persistence, meta_parameters)
    , _cuda_device(cuda_device)
  {}

  std::pair<tf::Task, tf::Task> build_execution_graph(util::Range<PointBuffer::PointIterator> points,
                                                      const AABB& bounds,
                                                      uint32_t num_indexing_threads,
                                                      tf::Taskflow& tf) override
  {
    // one task, like the single-threaded sort task of V1 (TilingAlgorithms.cpp:600-604); the GPU does
    // index + sort + every sampling level inside it
    auto task = tf.emplace([this, points, bounds, num_indexing_threads]() mutable {
      if (_batches_done++)
        throw std::runtime_error{ "TilingAlgorithmGPU: only single-batch runs are supported; raise "
                                  "--internal-cache-size to the number of points" };
      _root_bounds = bounds;
      const double bmin[3] = { bounds.min.x, bounds.min.y, bounds.min.z };
      const double bmax[3] = { bounds.max.x, bounds.max.y, bounds.max.z };
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
      _points.emplace(points);
      const auto n = static_cast<uint64_t>(points.size());
      // PointBuffer::positions() is a std::vector<Vector3<double>>: AoS x,y,z doubles (PointBuffer.h:291)
      double* xyz = n ? &(*std::begin(points)).position().x : nullptr;
      _tiler->index_batch(xyz, n); // clamps outliers in place, as index_point does
      // ACCURATE is complete here; FAST still owes the reconstructed upper levels (finalize)
      if (_meta_parameters.tiling_strategy != TilingStrategy::Fast)
        hand_off(/*count_progress=*/true);
    });
    return { task, task };
  }

  void finalize(const AABB& bounds) override
  {
    if (!_tiler || _meta_parameters.tiling_strategy != TilingStrategy::Fast)
      return;
    _tiler->finalize();
    hand_off(/*count_progress=*/true);
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
  void hand_off(bool count_progress)
  {
    const auto result = _tiler->result();
    const double rmin[3] = { _root_bounds.min.x, _root_bounds.min.y, _root_bounds.min.z };
    const double rmax[3] = { _root_bounds.max.x, _root_bounds.max.y, _root_bounds.max.z };
    std::vector<PointBuffer::PointReference> refs;
    const auto first_point = std::begin(*_points);
    for (const sw_node& node : result.nodes) {
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
  }

  int _cuda_device;
  size_t _batches_done = 0;
  AABB _root_bounds;
  std::optional<util::Range<PointBuffer::PointIterator>> _points; // PointIterator has no default ctor
  std::unique_ptr<swgpu::Tiler> _tiler;
};
