// swgpu_tiler.hpp — header-only C++17 RAII layer over the C ABI (include/swgpu.h).
//
// This is the host-side mirror of the reference's tiler interfaces for code that is compiled
// C++ (the reference itself): strategy names as the CLI spells them
// (core/tiling/Sampling.h:774-791, core/process/Tiler.cpp:189-198), TilerMetaParameters-shaped
// construction (core/process/Tiler.h:64-75), build_execution_graph / finalize
// (core/tiling/TilingAlgorithms.h:70-116), Potree node names (TilingAlgorithms.cpp:139),
// node bounds by the halving recurrence (core/tiling/OctreeAlgorithms.cpp:3-18) and the reference's
// error behaviour: every failure is a std::runtime_error carrying the library's message.
// It depends on nothing but the C ABI; TilingAlgorithmGPU.h (next to this file) plugs it into the
// reference's TilingAlgorithmBase.
#pragma once

#include "../../include/swgpu.h"

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace swgpu {

struct Error : std::runtime_error
{
  Error(int code_, const std::string& what)
    : std::runtime_error(what)
    , code(code_)
  {}
  int code;
};

inline sw_sampling
sampling_from_name(const std::string& name)
{
  if (name == "RANDOM_GRID")
    return SW_RANDOM_GRID;
  if (name == "GRID_CENTER")
    return SW_GRID_CENTER;
  if (name == "MIN_DISTANCE")
    return SW_MIN_DISTANCE;
  if (name == "JITTERED")
    return SW_JITTERED;
  if (name == "MIN_DISTANCE_FAST")
    return SW_MIN_DISTANCE_FAST;
  throw Error(SW_ERR_INVALID_ARGUMENT, "Unrecognized sampling strategy " + name);
}

inline sw_tiling
tiling_from_name(const std::string& name)
{
  if (name == "ACCURATE")
    return SW_ACCURATE;
  if (name == "FAST")
    return SW_FAST;
  throw Error(SW_ERR_INVALID_ARGUMENT, "Unrecognized tiling strategy " + name);
}

// "r" + one octant digit per level, octant = x<<2 | y<<1 | z (TilingAlgorithms.cpp:139)
inline std::string
node_name(uint64_t index, uint32_t levels)
{
  std::string name = "r";
  for (uint32_t l = 0; l < levels; ++l)
    name.push_back(static_cast<char>('0' + ((index >> (3 * (levels - 1 - l))) & 7)));
  return name;
}

// get_bounds_from_node_index: iterate get_octant_bounds from the root (OctreeAlgorithms.cpp:3-18,
// 64-72); the recurrence, not a closed form, so the doubles match the reference bit for bit.
inline void
node_bounds(uint64_t index, uint32_t levels, const double root_min[3], const double root_max[3], double out_min[3],
            double out_max[3])
{
  for (int a = 0; a < 3; ++a) {
    out_min[a] = root_min[a];
    out_max[a] = root_max[a];
  }
  for (uint32_t l = 0; l < levels; ++l) {
    const uint32_t octant = static_cast<uint32_t>((index >> (3 * (levels - 1 - l))) & 7);
    const uint32_t bit[3] = { (octant >> 2) & 1u, (octant >> 1) & 1u, octant & 1u };
    for (int a = 0; a < 3; ++a) {
      const double half = (out_max[a] - out_min[a]) / 2;
      if (bit[a])
        out_min[a] = out_min[a] + half;
      out_max[a] = out_min[a] + half;
    }
  }
}

class Tiler
{
public:
  Tiler(sw_sampling sampling, sw_tiling tiling, float spacing_at_root, uint32_t max_depth, uint64_t max_points_per_node,
        const double bounds_min[3], const double bounds_max[3], uint32_t num_indexing_threads, int device = 0)
  {
    _params.sampling = sampling;
    _params.tiling = tiling;
    _params.spacing_at_root = spacing_at_root;
    _params.max_depth = max_depth;
    _params.max_points_per_node = max_points_per_node;
    for (int a = 0; a < 3; ++a) {
      _params.bounds_min[a] = bounds_min[a];
      _params.bounds_max[a] = bounds_max[a];
    }
    _params.concurrency = num_indexing_threads;
    _params.reserved = 0;
    const int rc = swgpu_create(&_params, device, &_handle);
    if (rc != SW_OK)
      throw Error(rc,
                  rc == SW_ERR_CUDA ? "swgpu_create: no usable CUDA device (there is no CPU fallback)"
                                    : "swgpu_create: invalid tiler parameters");
  }
  ~Tiler() { swgpu_destroy(_handle); }
  Tiler(const Tiler&) = delete;
  Tiler& operator=(const Tiler&) = delete;

  // build_execution_graph for one batch: xyz = PointBuffer::positions() (AoS doubles); outliers are
  // clamped in place exactly as index_point does (OctreeAlgorithms.h:156-170)
  void index_batch(double* xyz_host, uint64_t n) { check(swgpu_index_batch(_handle, xyz_host, n)); }
  // The same batch while it is still in LAS record form (laszip_point::X/Y/Z): the device computes what
  // position_from_las_point (core/io/LASFile.cpp:79-94) and the tiler's shift-to-centre transformation
  // (core/process/TilerProcess.cpp:552-559) compute on the host, fused with the indexing kernel.
  void index_batch_las(const int32_t* las_xyz_host, uint64_t n, const sw_las_transform& transform)
  {
    check(swgpu_index_batch_las(_handle, las_xyz_host, n, &transform));
  }
  // PointBuffer positions of the last batch (n x 3 doubles, original order, after index_point's clamping)
  void positions(double* xyz_host) { check(swgpu_get_positions(_handle, xyz_host)); }
  void finalize() { check(swgpu_finalize(_handle)); }
  // Multi-batch mode: every later index_batch is one build_execution_graph against what the earlier batches
  // stored (tile_node with cached points, core/tiling/TilingAlgorithms.cpp:351-492); result() then returns the
  // final content of every node with GLOBAL point ids (points of earlier batches + index in the batch).
  void set_multi_batch(bool on) { check(swgpu_set_multi_batch(_handle, on ? 1 : 0)); }
  // Nodes where the reference would re-root (TilingAlgorithms.cpp:444-483): fail (default) or store them whole,
  // flagged SW_NODE_TERMINAL | SW_NODE_DEEP
  void set_deep_node_policy(bool store_whole) { check(swgpu_set_deep_node_policy(_handle, store_whole ? 1 : 0)); }
  // K2: -1 automatic (top-digit passes + segment finish), 0 eight LSD passes, 1..3 explicit; same order in every mode
  void set_sort_mode(int mode) { check(swgpu_set_sort_mode(_handle, mode)); }

  struct Result
  {
    std::vector<sw_node> nodes;
    std::vector<uint32_t> point_ids; // node-major, Morton order inside a node
  };

  Result result()
  {
    uint64_t n_nodes = 0, n_ids = 0;
    check(swgpu_result_size(_handle, &n_nodes, &n_ids));
    Result r;
    r.nodes.resize(n_nodes);
    r.point_ids.resize(n_ids);
    check(swgpu_get_nodes(_handle, r.nodes.data(), r.point_ids.data()));
    return r;
  }

  // Position payloads of ALL nodes in the node-major order of result(): what PNTSWriter's PositionAttribute
  // (core/io/PNTSWriter.cpp:326-342) and LASPersistence::persist_points (core/io/LASPersistence.h:119-131,
  // 160-163) store.  Row i of the node table owns records [first, first + count).
  std::vector<float> payload_pnts()
  {
    uint64_t n_nodes = 0, n_ids = 0;
    check(swgpu_result_size(_handle, &n_nodes, &n_ids));
    std::vector<float> out(3 * n_ids);
    if (n_ids)
      check(swgpu_get_payload_pnts(_handle, out.data()));
    return out;
  }
  struct LasPayload
  {
    std::vector<int32_t> records;            // n_point_ids x 3
    std::vector<sw_las_node_header> headers; // one per node table row
  };
  LasPayload payload_las()
  {
    uint64_t n_nodes = 0, n_ids = 0;
    check(swgpu_result_size(_handle, &n_nodes, &n_ids));
    LasPayload p;
    p.records.resize(3 * n_ids);
    p.headers.resize(n_nodes);
    if (n_ids)
      check(swgpu_get_payload_las(_handle, p.records.data(), p.headers.data()));
    return p;
  }

  int32_t start_level()
  {
    int32_t s = -1;
    check(swgpu_get_start_level(_handle, &s));
    return s;
  }

  const sw_params& params() const { return _params; }
  swgpu_handle handle() const { return _handle; }

private:
  void check(int rc)
  {
    if (rc != SW_OK)
      throw Error(rc, swgpu_last_error(_handle));
  }

  sw_params _params{};
  swgpu_handle _handle = nullptr;
};

// Several GPUs from the one process the reference is (core/process/Tiler.cpp:189-198): swgpu_multi_* cuts the
// batch into slices, shuffles the points over NVLink so that every GPU owns whole Morton-prefix subtrees, tiles
// the shards on one host thread per GPU and merges the node tables.  Same result type as Tiler.
class MultiTiler
{
public:
  MultiTiler(sw_sampling sampling, sw_tiling tiling, float spacing_at_root, uint32_t max_depth,
             uint64_t max_points_per_node, const double bounds_min[3], const double bounds_max[3],
             uint32_t num_indexing_threads, const std::vector<int>& devices)
  {
    _params.sampling = sampling;
    _params.tiling = tiling;
    _params.spacing_at_root = spacing_at_root;
    _params.max_depth = max_depth;
    _params.max_points_per_node = max_points_per_node;
    for (int a = 0; a < 3; ++a) {
      _params.bounds_min[a] = bounds_min[a];
      _params.bounds_max[a] = bounds_max[a];
    }
    _params.concurrency = num_indexing_threads;
    _params.reserved = 0;
    const int rc = swgpu_multi_create(&_params, devices.data(), static_cast<uint32_t>(devices.size()), &_handle);
    if (rc != SW_OK)
      throw Error(rc,
                  rc == SW_ERR_CUDA ? "swgpu_multi_create: a CUDA device is missing or the GPUs cannot access each other"
                                    : "swgpu_multi_create: invalid tiler parameters");
  }
  ~MultiTiler() { swgpu_multi_destroy(_handle); }
  MultiTiler(const MultiTiler&) = delete;
  MultiTiler& operator=(const MultiTiler&) = delete;

  void index_batch(double* xyz_host, uint64_t n) { check(swgpu_multi_index_batch(_handle, xyz_host, n, nullptr, 0)); }
  void finalize() { check(swgpu_multi_finalize(_handle)); }
  Tiler::Result result()
  {
    uint64_t n_nodes = 0, n_ids = 0;
    check(swgpu_multi_result_size(_handle, &n_nodes, &n_ids));
    Tiler::Result r;
    r.nodes.resize(n_nodes);
    r.point_ids.resize(n_ids);
    check(swgpu_multi_get_nodes(_handle, r.nodes.data(), r.point_ids.data()));
    return r;
  }

private:
  void check(int rc)
  {
    if (rc != SW_OK)
      throw Error(rc, swgpu_multi_last_error(_handle));
  }
  sw_params _params{};
  swgpu_multi_handle _handle = nullptr;
};

} // namespace swgpu
