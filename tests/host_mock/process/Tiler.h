// stand-in for core/process/Tiler.h:30-75 (TilingStrategy, thread counts, TilerMetaParameters)
#pragma once
#include <cstddef>
#include <cstdint>
#include <variant>
enum class TilingStrategy { Accurate, Fast };
struct FixedThreadCount { uint32_t num_threads_for_reading; uint32_t num_threads_for_indexing; };
struct AdaptiveThreadCount { uint32_t num_threads; };
struct TilerMetaParameters
{
  float spacing_at_root;
  uint32_t max_depth;
  size_t max_points_per_node;
  size_t batch_read_size;
  size_t internal_cache_size;
  bool shift_points_to_origin;
  bool create_journal;
  TilingStrategy tiling_strategy;
  std::variant<FixedThreadCount, AdaptiveThreadCount> thread_count;
};
