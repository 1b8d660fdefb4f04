/*
 * oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's OWN hot-path code, compiled verbatim from /root/reference (never copied
 * into this repository), behind the same C API shape as tiler_oracle.cpp (prefix swr_ instead of
 * swo_).  Built by oracle/Makefile into oracle/_ref/libswref.so, only where /root/reference
 * exists; the built library travels to the GPU box, the sources do not.
 *
 * Reference functions called (all from /root/reference/schwarzwald/core):
 *   index_point<21>                         tiling/OctreeAlgorithms.h:145-175
 *   partition_points_into_child_octants     tiling/OctreeAlgorithms.h:240-265
 *   get_octant_bounds                       tiling/OctreeAlgorithms.cpp:3-18
 *   get_bounds_from_morton_index<21>        tiling/OctreeAlgorithms.h:104-116
 *   sample_points(SamplingStrategy&, ...)   tiling/Sampling.h:799-821 (all four strategies)
 *   required_morton_index_depth             tiling/Sampling.cpp:29-62
 *   expand_bits_by_3 / contract_bits_by_3   util/stuff.h:207-234
 *   position_from_las_point                 io/LASFile.cpp:79-94
 *   attributes::PositionAttribute           io/PNTSWriter.cpp:326-342,344-360
 *   LASPersistence::persist_points<Iter>    io/LASPersistence.h:33-240 (+ compute_las_scale_from_bounds,
 *                                           io/LASPersistence.cpp:17-28)
 * TilingAlgorithms.cpp needs taskflow/boost::hana/cista, which are not in this image, so the
 * control flow around these calls is the restatement in orchestrator.h.
 */
#include "orchestrator.h"

#include "datastructures/PointBuffer.h"
#include "tiling/Node.h"
#include "tiling/OctreeAlgorithms.h"
#include "tiling/Sampling.h"
#include "util/stuff.h"

/* the reader's position conversion and the writers' position payloads,
 * compiled verbatim as well (io/LASFile.cpp, io/LASPersistence.cpp, io/PNTSWriter.cpp) against the
 * in-memory LASzip stand-in oracle/shim/laszip_api.h */
#include "io/LASFile.h"
#include "io/LASPersistence.h"
#include "io/PNTSWriter.h"

#include <chrono>
#include <cstring>
#include <memory>

/* util/Transformation.cpp needs PROJ and is not compiled; io/PNTSWriter.cpp references this one function
 * from it in writePNTSFile (converter mode), which nothing here calls. */
Vector3<double>
setOriginToSmallestPoint(std::vector<Vector3<double>>&)
{
  throw std::logic_error("setOriginToSmallestPoint is not part of the tiler hot path");
}

namespace {

AABB
to_aabb(const swo::Box& b)
{
  return AABB{ { b.min[0], b.min[1], b.min[2] }, { b.max[0], b.max[1], b.max[2] } };
}

swo::Box
from_aabb(const AABB& a)
{
  swo::Box b;
  b.min[0] = a.min.x;
  b.min[1] = a.min.y;
  b.min[2] = a.min.z;
  b.max[0] = a.max.x;
  b.max[1] = a.max.y;
  b.max[2] = a.max.z;
  return b;
}

SamplingStrategy
make_strategy(int32_t sampling, size_t max_points)
{
  switch (sampling) {
    case SW_RANDOM_GRID:
      return RandomSortedGridSampling{ max_points };
    case SW_GRID_CENTER:
      return GridCenterSampling{ max_points };
    case SW_MIN_DISTANCE:
      return PoissonDiskSampling{ max_points };
    case SW_JITTERED:
      return JitteredSampling{ max_points };
    case SW_MIN_DISTANCE_FAST: /* the strategy object TilerProcess::make_sampling_strategy builds for
                                  "MIN_DISTANCE_FAST" (process/TilerProcess.cpp:500-508; TilerProcess.cpp
                                  itself needs boost/taskflow and is not compiled here) */
      return AdaptivePoissonDiskSampling{ max_points, [](int32_t node_level) -> float {
                                           if (node_level < 0)
                                             return 0.25f;
                                           if (node_level < 1)
                                             return 0.5f;
                                           return 1.f;
                                         } };
  }
  throw swo::OracleError(SW_ERR_INVALID_ARGUMENT, "unknown sampling strategy");
}

struct RefPrims
{
  using Item = IndexedPoint64;

  PointBuffer buffer; /* owns a copy of the positions; PointReference indexes into it */
  SamplingStrategy strategy;
  std::vector<PointBuffer::PointReference> refs;

  RefPrims(const double* xyz, uint64_t n, int32_t sampling, size_t max_points)
    : strategy(make_strategy(sampling, max_points))
  {
    std::vector<Vector3<double>> positions(n);
    for (uint64_t i = 0; i < n; ++i)
      positions[i] = { xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] };
    buffer = PointBuffer(n, std::move(positions));
    refs.reserve(n);
    for (auto it = buffer.begin(); it != buffer.end(); ++it)
      refs.push_back(*it);
  }

  uint64_t key(const Item& i) const { return i.morton_index.get(); }
  uint32_t id(const Item& i) const
  {
    return static_cast<uint32_t>(&i.point_reference.position() - buffer.positions().data());
  }

  void index_all(std::vector<Item>& out, const swo::Box& bounds)
  {
    const AABB b = to_aabb(bounds);
    out.clear();
    out.reserve(refs.size());
    for (auto& r : refs)
      out.push_back(index_point<21>(r, b, OutlierPointsBehaviour::ClampToBounds));
  }

  uint64_t point_count() const { return refs.size(); }

  /* the reference's own calculate_morton_index<21> (tiling/OctreeAlgorithms.h:64-87) relative to arbitrary
   * bounds, as read_pnts_from_disk calls it (tiling/TilingAlgorithms.cpp:88-90) */
  uint64_t morton_in_bounds(uint32_t id, const swo::Box& bounds)
  {
    return calculate_morton_index<21>(refs[id].position(), to_aabb(bounds)).get();
  }
  Item make_item(uint32_t id, uint64_t key) { return Item{ refs[id], MortonIndex64{ key } }; }

  void index_range(uint64_t b, uint64_t e, std::vector<Item>& out, const swo::Box& bounds)
  {
    const AABB bb = to_aabb(bounds);
    out.clear();
    out.reserve(e - b);
    for (uint64_t i = b; i < e; ++i)
      out.push_back(index_point<21>(refs[i], bb, OutlierPointsBehaviour::ClampToBounds));
  }

  void index_ids(const std::vector<uint32_t>& ids, std::vector<Item>& out, const swo::Box& bounds)
  {
    const AABB b = to_aabb(bounds);
    out.clear();
    out.reserve(ids.size());
    for (uint32_t id : ids)
      out.push_back(index_point<21>(refs[id], b, OutlierPointsBehaviour::ClampToBounds));
  }

  swo::Box octant_bounds(uint8_t octant, const swo::Box& parent) const
  {
    return from_aabb(get_octant_bounds(octant, to_aabb(parent)));
  }

  std::array<size_t, 9> partition(const Item* begin, const Item* end, uint32_t level) const
  {
    const auto parts = partition_points_into_child_octants(begin, end, level);
    std::array<size_t, 9> cuts{};
    cuts[0] = 0;
    for (int o = 0; o < 8; ++o)
      cuts[o + 1] = static_cast<size_t>(parts[o].end() - begin);
    return cuts;
  }

  int32_t required_depth(int32_t node_level, const swo::NodeStructure& root) const
  {
    octree::NodeStructure r;
    r.bounds = to_aabb(root.bounds);
    r.level = root.level;
    r.max_spacing = root.max_spacing;
    r.max_depth = root.max_depth;
    r.morton_index = {};
    r.name = "r";
    return required_morton_index_depth(strategy, node_level, r);
  }

  size_t sample(Item* begin,
                Item* end,
                uint64_t node_key,
                int32_t node_level,
                const swo::Box& root_bounds,
                float spacing_at_root,
                swo::Behaviour behaviour)
  {
    try {
      Item* pp = sample_points(strategy,
                               begin,
                               end,
                               MortonIndex64{ node_key },
                               node_level,
                               to_aabb(root_bounds),
                               spacing_at_root,
                               behaviour == swo::AlwaysAdhereToMinSpacing
                                 ? SamplingBehaviour::AlwaysAdhereToMinSpacing
                                 : SamplingBehaviour::TakeAllWhenCountBelowMaxPoints);
      return static_cast<size_t>(pp - begin);
    } catch (const std::runtime_error& e) {
      const std::string what = e.what();
      if (what.find("smaller than 16x16") != std::string::npos)
        throw swo::OracleError(SW_ERR_JITTER_GRID_TOO_SMALL, what);
      if (what.find("too small to be sampled") != std::string::npos)
        throw swo::OracleError(SW_ERR_JITTER_NODE_TOO_SMALL, what);
      throw;
    }
  }
};

struct Handle
{
  std::vector<sw_node> nodes;
  std::vector<uint32_t> ids;
  std::vector<uint64_t> keys;
  std::vector<uint32_t> order;
  uint64_t duplicate_keys = 0;
  int32_t start_level = -1;
  double seconds = 0.0; /* index + sort + tiling only (o.run()), without building or copying buffers */
  std::string error;
};

swo::Box
make_box(const double* bmin, const double* bmax)
{
  swo::Box b;
  for (int a = 0; a < 3; ++a) {
    b.min[a] = bmin[a];
    b.max[a] = bmax[a];
  }
  return b;
}

} // namespace

extern "C" {

uint64_t
swr_expand_bits_by_3(uint64_t v)
{
  return expand_bits_by_3(v);
}

uint64_t
swr_contract_bits_by_3(uint64_t v)
{
  return contract_bits_by_3(v);
}

void
swr_index_points(double* xyz, uint64_t n, const double* bmin, const double* bmax, uint64_t* keys)
{
  RefPrims p(xyz, n, SW_RANDOM_GRID, 1);
  std::vector<IndexedPoint64> items;
  p.index_all(items, make_box(bmin, bmax));
  for (uint64_t i = 0; i < n; ++i) {
    keys[i] = items[i].morton_index.get();
    const auto& pos = p.buffer.positions()[i]; /* index_point clamps in place */
    xyz[3 * i] = pos.x;
    xyz[3 * i + 1] = pos.y;
    xyz[3 * i + 2] = pos.z;
  }
}

void
swr_octant_bounds(uint8_t octant, const double* bmin, const double* bmax, double* out6)
{
  const swo::Box r = from_aabb(get_octant_bounds(octant, to_aabb(make_box(bmin, bmax))));
  std::memcpy(out6, r.min, 3 * sizeof(double));
  std::memcpy(out6 + 3, r.max, 3 * sizeof(double));
}

void
swr_bounds_from_morton_index(uint64_t key, uint32_t depth, const double* bmin, const double* bmax, double* out6)
{
  const swo::Box r =
    from_aabb(get_bounds_from_morton_index(MortonIndex64{ key }, to_aabb(make_box(bmin, bmax)), depth));
  std::memcpy(out6, r.min, 3 * sizeof(double));
  std::memcpy(out6 + 3, r.max, 3 * sizeof(double));
}

int32_t
swr_required_morton_index_depth(int32_t sampling,
                                int32_t node_level,
                                const double* bmin,
                                const double* bmax,
                                float root_max_spacing)
{
  double dummy[3] = { 0, 0, 0 };
  RefPrims p(dummy, 1, sampling, 1);
  swo::NodeStructure root{};
  root.bounds = make_box(bmin, bmax);
  root.max_spacing = root_max_spacing;
  root.level = -1;
  return p.required_depth(node_level, root);
}

void
swr_partition_child_octants(const uint64_t* keys, uint64_t n, uint32_t level, uint64_t* cuts9)
{
  std::vector<IndexedPoint64> items(n);
  for (uint64_t i = 0; i < n; ++i)
    items[i].morton_index = MortonIndex64{ keys[i] };
  const auto parts = partition_points_into_child_octants(items.data(), items.data() + n, level);
  cuts9[0] = 0;
  for (int o = 0; o < 8; ++o)
    cuts9[o + 1] = static_cast<uint64_t>(parts[o].end() - items.data());
}

int64_t
swr_sample_points(int32_t sampling,
                  const double* xyz,
                  const uint64_t* keys,
                  const uint32_t* ids,
                  uint64_t n,
                  uint64_t node_key,
                  int32_t node_level,
                  const double* bmin,
                  const double* bmax,
                  float spacing_at_root,
                  int32_t behaviour,
                  uint64_t max_points_per_node,
                  uint64_t* keys_out,
                  uint32_t* ids_out)
{
  try {
    uint32_t max_id = 0;
    for (uint64_t i = 0; i < n; ++i)
      max_id = std::max(max_id, ids[i]);
    RefPrims p(xyz, n ? static_cast<uint64_t>(max_id) + 1 : 0, sampling, max_points_per_node);
    std::vector<IndexedPoint64> items(n);
    for (uint64_t i = 0; i < n; ++i) {
      items[i].point_reference = p.refs[ids[i]];
      items[i].morton_index = MortonIndex64{ keys[i] };
    }
    const size_t taken = p.sample(items.data(),
                                  items.data() + n,
                                  node_key,
                                  node_level,
                                  make_box(bmin, bmax),
                                  spacing_at_root,
                                  static_cast<swo::Behaviour>(behaviour));
    for (uint64_t i = 0; i < n; ++i) {
      keys_out[i] = items[i].morton_index.get();
      ids_out[i] = p.id(items[i]);
    }
    return static_cast<int64_t>(taken);
  } catch (const swo::OracleError& e) {
    return -static_cast<int64_t>(e.code);
  }
}

/* worker threads of the following swr_tile calls: the reference's taskflow workers (one task per start
 * node / child subtree) become std::threads over the same units of work, see orchestrator.h */
static unsigned g_threads = 1;

void
swr_set_threads(uint32_t n)
{
  g_threads = n ? n : 1;
}

/* 1 = sort with std::sort as the reference does (timing runs), 0 = std::stable_sort (parity runs, default) */
static int g_reference_sort = 0;

void
swr_set_reference_sort(int32_t on)
{
  g_reference_sort = on;
}

/* >= 0: FAST tiles with this start level instead of estimating it (subtree parity checks), -1 = estimate */
static int32_t g_start_level_override = -1;

void
swr_set_start_level_override(int32_t level)
{
  g_start_level_override = level;
}

/* `passes` runs of the whole path over the same PointBuffer (built once: constructing it is not part of the hot
 * path); seconds_out[k] = index + sort + tiling time of pass k, the handle holds the result of the last pass.
 * index_point's in-place clamp is idempotent, so every pass does the same work. */
int
swr_tile_repeat(const sw_params* params, double* xyz, uint64_t n, uint32_t passes, double* seconds_out,
                void** out_handle)
{
  auto* h = new Handle();
  *out_handle = h;
  try {
    RefPrims prims(xyz, n, params->sampling, params->max_points_per_node);
    for (uint32_t k = 0; k < (passes ? passes : 1u); ++k) {
      swo::Orchestrator<RefPrims> o(prims, *params, g_threads);
      o.reference_sort = g_reference_sort != 0;
      o.start_level_override = g_start_level_override;
      const auto t0 = std::chrono::steady_clock::now();
      o.run();
      h->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (seconds_out)
        seconds_out[k] = h->seconds;
      h->nodes = std::move(o.nodes);
      h->ids = std::move(o.ids);
      h->keys = std::move(o.sorted_keys);
      h->order = std::move(o.sorted_ids);
      h->duplicate_keys = o.duplicate_keys;
      h->start_level = o.start_level;
    }
    /* index_point clamps in place inside the PointBuffer: hand the clamped positions back */
    for (uint64_t i = 0; i < n; ++i) {
      const auto& pos = prims.buffer.positions()[i];
      xyz[3 * i] = pos.x;
      xyz[3 * i + 1] = pos.y;
      xyz[3 * i + 2] = pos.z;
    }
    return SW_OK;
  } catch (const swo::OracleError& e) {
    h->error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->error = e.what();
    return SW_ERR_STATE;
  }
}

int
swr_tile(const sw_params* params, double* xyz, uint64_t n, void** out_handle)
{
  return swr_tile_repeat(params, xyz, n, 1, nullptr, out_handle);
}

double
swr_tile_seconds(void* handle)
{
  return static_cast<Handle*>(handle)->seconds;
}

int
swr_tile_batches(const sw_params* params, double* xyz, uint64_t n, const uint64_t* offsets, uint32_t n_batches,
                 void** out_handle)
{
  auto* h = new Handle();
  *out_handle = h;
  try {
    if (n_batches == 0 || offsets[n_batches] != n)
      throw swo::OracleError(SW_ERR_INVALID_ARGUMENT, "tile_batches: offsets must end at n");
    RefPrims prims(xyz, n, params->sampling, params->max_points_per_node);
    swo::Orchestrator<RefPrims> o(prims, *params, 1);
    if (params->tiling == SW_FAST)
      o.run_fast_batches(offsets, n_batches);
    else
      o.run_accurate_batches(offsets, n_batches);
    h->start_level = o.start_level;
    h->nodes = std::move(o.nodes);
    h->ids = std::move(o.ids);
    h->keys = std::move(o.sorted_keys);
    h->order = std::move(o.sorted_ids);
    for (uint64_t i = 0; i < n; ++i) { /* index_point clamps in place inside the PointBuffer */
      const auto& pos = prims.buffer.positions()[i];
      xyz[3 * i] = pos.x;
      xyz[3 * i + 1] = pos.y;
      xyz[3 * i + 2] = pos.z;
    }
    return SW_OK;
  } catch (const swo::OracleError& e) {
    h->error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->error = e.what();
    return SW_ERR_STATE;
  }
}

uint64_t
swr_node_count(void* handle)
{
  return static_cast<Handle*>(handle)->nodes.size();
}

uint64_t
swr_point_id_count(void* handle)
{
  return static_cast<Handle*>(handle)->ids.size();
}

int32_t
swr_start_level(void* handle)
{
  return static_cast<Handle*>(handle)->start_level;
}

uint64_t
swr_duplicate_keys(void* handle)
{
  return static_cast<Handle*>(handle)->duplicate_keys;
}

void
swr_get_nodes(void* handle, sw_node* nodes, uint32_t* ids)
{
  auto* h = static_cast<Handle*>(handle);
  if (nodes)
    std::memcpy(nodes, h->nodes.data(), h->nodes.size() * sizeof(sw_node));
  if (ids)
    std::memcpy(ids, h->ids.data(), h->ids.size() * sizeof(uint32_t));
}

void
swr_get_keys(void* handle, uint64_t* keys, uint32_t* order)
{
  auto* h = static_cast<Handle*>(handle);
  if (keys)
    std::memcpy(keys, h->keys.data(), h->keys.size() * sizeof(uint64_t));
  if (order)
    std::memcpy(order, h->order.data(), h->order.size() * sizeof(uint32_t));
}

const char*
swr_last_error(void* handle)
{
  return static_cast<Handle*>(handle)->error.c_str();
}

void
swr_destroy(void* handle)
{
  delete static_cast<Handle*>(handle);
}

/* ---- LAS input transform and writer payloads ------------------------- */

void
swr_las_positions(const int32_t* las, uint64_t n, const sw_las_transform* t, double* xyz)
{
  laszip_header header;
  header.x_scale_factor = t->scale[0];
  header.y_scale_factor = t->scale[1];
  header.z_scale_factor = t->scale[2];
  header.x_offset = t->offset[0];
  header.y_offset = t->offset[1];
  header.z_offset = t->offset[2];
  header.min_x = t->header_min[0];
  header.min_y = t->header_min[1];
  header.min_z = t->header_min[2];
  header.max_x = t->header_max[0];
  header.max_y = t->header_max[1];
  header.max_z = t->header_max[2];
  const Vector3<double> center{ t->center[0], t->center[1], t->center[2] };
  for (uint64_t i = 0; i < n; ++i) {
    laszip_point point;
    point.X = las[3 * i];
    point.Y = las[3 * i + 1];
    point.Z = las[3 * i + 2];
    auto position = position_from_las_point(point, header); /* the reference's own function */
    if (t->shift_to_center) {
      /* the transformation TilerProcess registers on the point source (process/TilerProcess.cpp:552-559;
       * that TU needs boost/taskflow and is not compiled here, so its four statements are spelled out) */
      position -= center;
      position.x = static_cast<float>(position.x);
      position.y = static_cast<float>(position.y);
      position.z = static_cast<float>(position.z);
    }
    xyz[3 * i] = position.x;
    xyz[3 * i + 1] = position.y;
    xyz[3 * i + 2] = position.z;
  }
}

void
swr_payload_pnts(const double* xyz, const uint32_t* ids, uint64_t n_ids, float* out)
{
  uint64_t n = 0;
  for (uint64_t j = 0; j < n_ids; ++j)
    n = std::max<uint64_t>(n, static_cast<uint64_t>(ids[j]) + 1);
  RefPrims p(xyz, n, SW_RANDOM_GRID, 1);
  std::vector<PointBuffer::PointReference> refs;
  refs.reserve(n_ids);
  for (uint64_t j = 0; j < n_ids; ++j)
    refs.push_back(p.refs[ids[j]]);
  attributes::PositionAttribute attribute;
  attribute.extractFromPoints(gsl::span<PointBuffer::PointReference>(refs.data(), refs.size()));
  const auto bytes = attribute.getBinaryDataRange();
  std::memcpy(out, bytes.data(), bytes.size());
}

void
swr_payload_las(const double* xyz, const uint32_t* ids, const sw_node* nodes, uint64_t n_nodes, const double* bmin,
                const double* bmax, int32_t* out, sw_las_node_header* headers)
{
  uint64_t n = 0, n_ids = 0;
  for (uint64_t r = 0; r < n_nodes; ++r)
    n_ids = std::max<uint64_t>(n_ids, nodes[r].first + nodes[r].count);
  for (uint64_t j = 0; j < n_ids; ++j)
    n = std::max<uint64_t>(n, static_cast<uint64_t>(ids[j]) + 1);
  RefPrims p(xyz, n, SW_RANDOM_GRID, 1);
  PointAttributes attributes;
  attributes.insert(PointAttribute::Position);
  LASPersistence sink("swr", attributes, attributes, Compressed::No);
  const AABB root = to_aabb(make_box(bmin, bmax));
  for (uint64_t r = 0; r < n_nodes; ++r) {
    AABB b = root;
    for (uint32_t l = 0; l < nodes[r].levels; ++l) /* get_bounds_from_node_index, OctreeAlgorithms.cpp:64-72 */
      b = get_octant_bounds(static_cast<uint8_t>((nodes[r].index >> (3 * (nodes[r].levels - 1 - l))) & 7), b);
    if (headers) {
      headers[r] = sw_las_node_header{};
      headers[r].scale = compute_las_scale_from_bounds(b);
    }
    if (!nodes[r].count)
      continue;
    std::vector<PointBuffer::PointReference> refs;
    for (uint64_t j = nodes[r].first; j < nodes[r].first + nodes[r].count; ++j)
      refs.push_back(p.refs[ids[j]]);
    const std::string name = "node" + std::to_string(r);
    sink.persist_points(refs.begin(), refs.end(), b, name);
    const auto it = laszip_shim::files().find("swr/" + name + ".las");
    if (it == laszip_shim::files().end() || it->second.points.size() != nodes[r].count)
      throw std::runtime_error("LASPersistence did not write node " + name);
    const auto& file = it->second;
    if (headers) { /* what persist_points stored in the LAS header */
      headers[r].offset[0] = file.header.x_offset;
      headers[r].offset[1] = file.header.y_offset;
      headers[r].offset[2] = file.header.z_offset;
      headers[r].max[0] = file.header.max_x;
      headers[r].max[1] = file.header.max_y;
      headers[r].max[2] = file.header.max_z;
      headers[r].scale = file.header.x_scale_factor;
    }
    for (uint64_t k = 0; k < nodes[r].count; ++k) {
      out[3 * (nodes[r].first + k)] = file.points[k].X;
      out[3 * (nodes[r].first + k) + 1] = file.points[k].Y;
      out[3 * (nodes[r].first + k) + 2] = file.points[k].Z;
    }
    laszip_shim::files().erase(it);
  }
}

} /* extern "C" */
