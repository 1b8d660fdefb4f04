/*
 * oracle/tiler_oracle.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle, "port").
 *
 * CPU restatement of Schwarzwald's tiler compute core for one batch.  It is the checker for the
 * CUDA path in schwarzwald_b200/csrc; it is never linked into, imported by, or called from the
 * product.  Parity pinning: this restatement is checked (tests/test_oracle_*.py) against
 *   (1) every known-answer vector the reference's own tests hold for this path
 *       (S/test/TestOctreeIndexing.cpp, TestMortonIndex.cpp, TestOctreeNodeIndex.cpp, see
 *       SURVEY.md §4), and
 *   (2) the reference's OWN primitives compiled verbatim from /root/reference into
 *       oracle/_ref/libswref.so (oracle/ref_driver.cpp + oracle/Makefile), on random inputs, and
 *       the golden fixtures under tests/golden/ generated from that library.
 *
 * Each function cites the reference lines it follows; paths are relative to
 * /root/reference/schwarzwald/.  The reference is built for baseline x86-64 (no -march, so no
 * FMA contraction); this file must be compiled with -ffp-contract=off.
 */
#include "orchestrator.h"

#include <chrono>
#include <cstring>
#include <unordered_map>

namespace swo {

/* expand_bits_by_3(uint64_t), core/util/stuff.h:207-221 */
static inline uint64_t
expand_bits_by_3(uint64_t val)
{
  val &= 0x1FFFFFull; /* 21 bits */
  val = (val | (val << 32)) & 0x00FF00000000FFFFull;
  val = (val | (val << 16)) & 0x00FF0000FF0000FFull;
  val = (val | (val << 8)) & 0xF00F00F00F00F00Full;
  val = (val | (val << 4)) & 0x30C30C30C30C30C3ull; /* written as octal 0303030303030303030303 */
  val = (val | (val << 2)) & 0x1249249249249249ull;
  return val;
}

/* contract_bits_by_3, core/util/stuff.h:223-234 */
static inline uint64_t
contract_bits_by_3(uint64_t val)
{
  val &= 0x1249249249249249ull;
  val = (val | (val >> 2)) & 0x30C30C30C30C30C3ull;
  val = (val | (val >> 4)) & 0xF00F00F00F00F00Full;
  val = (val | (val >> 8)) & 0x00FF0000FF0000FFull;
  val = (val | (val >> 16)) & 0x00FF00000000FFFFull;
  val = (val | (val >> 32)) & 0x00000000FFFFFFFFull;
  return val;
}

/* get_prev_power_of_two(uint32_t), core/util/stuff.cpp:340-349 */
static inline uint32_t
get_prev_power_of_two(uint32_t x)
{
  x = x | (x >> 1);
  x = x | (x >> 2);
  x = x | (x >> 4);
  x = x | (x >> 8);
  x = x | (x >> 16);
  return x - (x >> 1);
}

/* calculate_morton_index<21>, core/tiling/OctreeAlgorithms.h:64-87 */
static inline uint64_t
calculate_morton_index(const double* p, const Box& b)
{
  const double sx = 2097152.0 / (b.max[0] - b.min[0]); /* std::pow(2, 21) / extent */
  const double sy = 2097152.0 / (b.max[1] - b.min[1]);
  const double sz = 2097152.0 / (b.max[2] - b.min[2]);
  const double nx = (p[0] - b.min[0]) * sx;
  const double ny = (p[1] - b.min[1]) * sy;
  const double nz = (p[2] - b.min[2]) * sz;
  const uint64_t cap = (1u << 21) - 1;
  const uint64_t bx = std::min(static_cast<uint64_t>(nx), cap);
  const uint64_t by = std::min(static_cast<uint64_t>(ny), cap);
  const uint64_t bz = std::min(static_cast<uint64_t>(nz), cap);
  return expand_bits_by_3(bz) | (expand_bits_by_3(by) << 1) | (expand_bits_by_3(bx) << 2);
}

/* index_point<21> with OutlierPointsBehaviour::ClampToBounds, OctreeAlgorithms.h:145-175.
 * AABB::isInside is inclusive (core/math/AABB.h:27-31); the clamp is written back. */
static inline uint64_t
index_point(double* p, const Box& b)
{
  const bool inside = p[0] >= b.min[0] && p[0] <= b.max[0] && p[1] >= b.min[1] &&
                      p[1] <= b.max[1] && p[2] >= b.min[2] && p[2] <= b.max[2];
  if (!inside) {
    p[0] = std::min(b.max[0], std::max(b.min[0], p[0]));
    p[1] = std::min(b.max[1], std::max(b.min[1], p[1]));
    p[2] = std::min(b.max[2], std::max(b.min[2], p[2]));
  }
  return calculate_morton_index(p, b);
}

/* get_octant_bounds, core/tiling/OctreeAlgorithms.cpp:3-18; octant = x<<2 | y<<1 | z */
static inline Box
get_octant_bounds(uint8_t octant, const Box& parent)
{
  const double ex = parent.max[0] - parent.min[0];
  const double ey = parent.max[1] - parent.min[1];
  const double ez = parent.max[2] - parent.min[2];
  Box r;
  r.min[2] = (octant & 1) ? (parent.min[2] + ez / 2) : parent.min[2];
  r.min[1] = ((octant >> 1) & 1) ? (parent.min[1] + ey / 2) : parent.min[1];
  r.min[0] = ((octant >> 2) & 1) ? (parent.min[0] + ex / 2) : parent.min[0];
  r.max[0] = r.min[0] + ex / 2;
  r.max[1] = r.min[1] + ey / 2;
  r.max[2] = r.min[2] + ez / 2;
  return r;
}

/* get_bounds_from_morton_index<21>, OctreeAlgorithms.h:104-116 */
static inline Box
get_bounds_from_morton_index(uint64_t key, const Box& root, uint32_t depth)
{
  Box b = root;
  const uint32_t max_level = std::min(depth, MAX_OCTREE_LEVELS);
  for (uint32_t level = 0; level < max_level; ++level) {
    const uint32_t shift = (MAX_OCTREE_LEVELS - level - 1) * 3;
    b = get_octant_bounds(static_cast<uint8_t>((key >> shift) & 7), b);
  }
  return b;
}

/* MortonIndex::truncate_to_level, core/datastructures/MortonIndex.h:123-129 (shift, not mask) */
static inline uint64_t
truncate_to_level(uint64_t key, uint32_t level)
{
  return key >> ((MAX_OCTREE_LEVELS - level - 1) * 3);
}

/* candidate level used INSIDE sample_points, Sampling.h:210-229 / 337-343: double spacing */
static inline int
candidate_level_in_sampler(const Box& root, float spacing_at_root, int32_t node_level)
{
  const double spacing_at_this_node = spacing_at_root / std::pow(2, node_level + 1);
  const float ratio = static_cast<float>((root.max[0] - root.min[0]) / spacing_at_this_node);
  return std::max(-1, static_cast<int>(std::floor(std::log2f(ratio))) - 1);
}

/* first_node_level_obeying_spacing + get_node_level_to_sample_from, core/tiling/Node.cpp:37-57:
 * the spacing is narrowed to a float PARAMETER before the division. */
static inline int32_t
node_level_to_sample_from(int32_t source_level, const NodeStructure& root)
{
  const double spacing_at_target_node = root.max_spacing / std::pow(2, source_level + 1);
  const float target_spacing = static_cast<float>(spacing_at_target_node);
  const float ratio = static_cast<float>((root.bounds.max[0] - root.bounds.min[0]) / target_spacing);
  return std::max(-1, static_cast<int>(std::floor(std::log2f(ratio))) - 1);
}

/* PERMUTATIONS_{16,32,64}, Sampling.h:14-138, regenerated into oracle/jitter_tables.inc by
 * tools/gen_jitter_tables.py (data tables, not code). */
#include "jitter_tables.inc"

static inline double
sqdist(const double* a, const double* b)
{
  /* Vector3::squaredDistanceTo -> (a-b).squaredLength() = x*x + y*y + z*z, math/Vector3.h:55-62 */
  const double dx = a[0] - b[0];
  const double dy = a[1] - b[1];
  const double dz = a[2] - b[2];
  return dx * dx + dy * dy + dz * dz;
}

struct IP
{
  uint64_t key;
  uint32_t id;
};

/* SparseGrid + GridCell, core/datastructures/SparseGrid.cpp:9-20,116-146, GridCell.cpp:10-58.
 * Cells link to every existing cell of their 3x3x3 block at creation (and are linked back), so
 * probing the 27 cell keys at query time visits exactly the same cells. */
struct SparseGridRestated
{
  int width, height, depth;
  Box aabb;
  float squaredSpacing;
  std::unordered_map<long long, std::vector<std::array<double, 3>>> cells;

  SparseGridRestated(const Box& b, float spacing)
    : aabb(b)
    , squaredSpacing(spacing * spacing)
  {
    const double cellSizeFactor = 5.0;
    width = static_cast<int>((b.max[0] - b.min[0]) / (spacing * cellSizeFactor));
    height = static_cast<int>((b.max[1] - b.min[1]) / (spacing * cellSizeFactor));
    depth = static_cast<int>((b.max[2] - b.min[2]) / (spacing * cellSizeFactor));
  }

  bool add(const double* p)
  {
    const double ex = aabb.max[0] - aabb.min[0];
    const double ey = aabb.max[1] - aabb.min[1];
    const double ez = aabb.max[2] - aabb.min[2];
    const int nx = static_cast<int>(width * (p[0] - aabb.min[0]) / ex);
    const int ny = static_cast<int>(height * (p[1] - aabb.min[1]) / ey);
    const int nz = static_cast<int>(depth * (p[2] - aabb.min[2]) / ez);
    const int i = std::max(0, std::min(nx, width - 1));
    const int j = std::max(0, std::min(ny, height - 1));
    const int k = std::max(0, std::min(nz, depth - 1));
    const double thr = squaredSpacing; /* GridCell::isDistant takes const double& */
    for (int ii = std::max(i - 1, 0); ii <= std::min(width - 1, i + 1); ++ii)
      for (int jj = std::max(j - 1, 0); jj <= std::min(height - 1, j + 1); ++jj)
        for (int kk = std::max(k - 1, 0); kk <= std::min(depth - 1, k + 1); ++kk) {
          const long long key = ((long long)kk << 40) | ((long long)jj << 20) | ii;
          auto it = cells.find(key);
          if (it == cells.end())
            continue;
          for (const auto& q : it->second)
            if (sqdist(p, q.data()) < thr)
              return false;
        }
    /* own cell when width/height/depth clamp produced an index outside the loops above */
    const long long own = ((long long)k << 40) | ((long long)j << 20) | i;
    auto& cell = cells[own];
    if (width <= 0 || height <= 0 || depth <= 0) {
      for (const auto& q : cell)
        if (sqdist(p, q.data()) < thr)
          return false;
    }
    cell.push_back({ p[0], p[1], p[2] });
    return true;
  }
};

struct RestatedPrims
{
  using Item = IP;

  double* xyz; /* AoS n x 3, clamped in place by index_all */
  uint64_t n;
  int32_t sampling;
  uint64_t max_points_per_node;

  static uint64_t key(const IP& i) { return i.key; }
  static uint32_t id(const IP& i) { return i.id; }

  void index_all(std::vector<IP>& out, const Box& bounds)
  {
    out.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
      out[i].key = index_point(xyz + 3 * i, bounds);
      out[i].id = static_cast<uint32_t>(i);
    }
  }

  uint64_t point_count() const { return n; }

  /* calculate_morton_index<21> relative to arbitrary bounds, no clamping (read_pnts_from_disk) */
  uint64_t morton_in_bounds(uint32_t id, const Box& bounds) const
  {
    return calculate_morton_index(xyz + 3 * static_cast<uint64_t>(id), bounds);
  }
  static IP make_item(uint32_t id, uint64_t key) { return IP{ key, id }; }

  void index_range(uint64_t b, uint64_t e, std::vector<IP>& out, const Box& bounds)
  {
    out.resize(e - b);
    for (uint64_t i = b; i < e; ++i) {
      out[i - b].key = index_point(xyz + 3 * i, bounds);
      out[i - b].id = static_cast<uint32_t>(i);
    }
  }

  void index_ids(const std::vector<uint32_t>& ids, std::vector<IP>& out, const Box& bounds)
  {
    out.resize(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {
      out[i].key = index_point(xyz + 3 * static_cast<uint64_t>(ids[i]), bounds);
      out[i].id = ids[i];
    }
  }

  Box octant_bounds(uint8_t octant, const Box& parent) const { return get_octant_bounds(octant, parent); }

  /* partition_points_into_child_octants, OctreeAlgorithms.h:240-265: returns the 9 cut offsets */
  std::array<size_t, 9> partition(const IP* begin, const IP* end, uint32_t level) const
  {
    std::array<size_t, 9> cuts{};
    const uint32_t shift = (MAX_OCTREE_LEVELS - level - 1) * 3;
    const IP* cur = begin;
    cuts[0] = 0;
    for (uint8_t octant = 0; octant < 8; ++octant) {
      while (cur != end && ((cur->key >> shift) & 7) <= octant)
        ++cur;
      cuts[octant + 1] = static_cast<size_t>(cur - begin);
    }
    return cuts;
  }

  /* required_morton_index_depth, core/tiling/Sampling.cpp:29-62 */
  int32_t required_depth(int32_t node_level, const NodeStructure& root) const
  {
    switch (sampling) {
      case SW_RANDOM_GRID:
      case SW_GRID_CENTER:
        return node_level_to_sample_from(node_level, root);
      case SW_MIN_DISTANCE:
      case SW_MIN_DISTANCE_FAST:
        return node_level;
      case SW_JITTERED: {
        const double spacing_at_this_node = root.max_spacing / std::pow(2, node_level + 1);
        const double perfect_cell_count =
          ((root.bounds.max[0] - root.bounds.min[0]) / std::pow(2, node_level + 1)) / spacing_at_this_node;
        const uint32_t actual_cell_count = get_prev_power_of_two(static_cast<uint32_t>(perfect_cell_count));
        const uint32_t levels = static_cast<uint32_t>(std::log2(actual_cell_count));
        return static_cast<int32_t>(static_cast<uint32_t>(node_level + levels));
      }
    }
    throw OracleError(SW_ERR_INVALID_ARGUMENT, "unknown sampling strategy");
  }

  /* stable_partition_with_jumps result layout, util/algorithms/Algorithm.h:22-77:
   * [selected in order | unselected in order] */
  static size_t stable_partition_flags(IP* begin, IP* end, const std::vector<uint8_t>& sel)
  {
    const size_t cnt = static_cast<size_t>(end - begin);
    std::vector<IP> tmp;
    tmp.reserve(cnt);
    for (size_t i = 0; i < cnt; ++i)
      if (sel[i])
        tmp.push_back(begin[i]);
    const size_t taken = tmp.size();
    for (size_t i = 0; i < cnt; ++i)
      if (!sel[i])
        tmp.push_back(begin[i]);
    std::copy(tmp.begin(), tmp.end(), begin);
    return taken;
  }

  size_t sample(IP* begin,
                IP* end,
                uint64_t node_key,
                int32_t node_level,
                const Box& root_bounds,
                float spacing_at_root,
                Behaviour behaviour)
  {
    const size_t cnt = static_cast<size_t>(end - begin);
    if (behaviour == TakeAllWhenCountBelowMaxPoints && cnt <= max_points_per_node)
      return cnt; /* Sampling.h:201-208, 328-335, 435-442, 612-619 */
    switch (sampling) {
      case SW_RANDOM_GRID:
        return sample_random_grid(begin, end, node_level, root_bounds, spacing_at_root);
      case SW_GRID_CENTER:
        return sample_grid_center(begin, end, node_level, root_bounds, spacing_at_root);
      case SW_MIN_DISTANCE:
        return sample_min_distance(begin, end, node_key, node_level, root_bounds, spacing_at_root);
      case SW_JITTERED:
        return sample_jittered(begin, end, node_key, node_level, root_bounds, spacing_at_root);
      case SW_MIN_DISTANCE_FAST:
        return sample_min_distance_fast(begin, end, node_key, node_level, root_bounds, spacing_at_root);
    }
    throw OracleError(SW_ERR_INVALID_ARGUMENT, "unknown sampling strategy");
  }

  /* RandomSortedGridSampling::sample_points, Sampling.h:187-308 */
  size_t sample_random_grid(IP* begin, IP* end, int32_t node_level, const Box& root, float spacing_at_root)
  {
    const size_t cnt = static_cast<size_t>(end - begin);
    const int cand = candidate_level_in_sampler(root, spacing_at_root, node_level);
    if (cand == -1) /* partition_at_root: take the first point */
      return cnt ? 1 : 0;
    std::vector<uint8_t> sel(cnt, 0);
    size_t cur = 0;
    while (cur < cnt) {
      sel[cur] = 1; /* first point of the run of equal truncate_to_level(cand) */
      const uint64_t cell = truncate_to_level(begin[cur].key, cand);
      size_t next = cur + 1;
      while (next < cnt && truncate_to_level(begin[next].key, cand) <= cell)
        ++next;
      cur = next;
    }
    return stable_partition_flags(begin, end, sel);
  }

  /* GridCenterSampling::sample_points, Sampling.h:314-416 */
  size_t sample_grid_center(IP* begin, IP* end, int32_t node_level, const Box& root, float spacing_at_root)
  {
    const size_t cnt = static_cast<size_t>(end - begin);
    const int cand = candidate_level_in_sampler(root, spacing_at_root, node_level);
    if (cand == -1)
      return 1; /* `return ++partition_point;` even for an empty range (Sampling.h:346-348) */
    std::vector<uint8_t> sel(cnt, 0);
    size_t cur = 0;
    while (cur < cnt) {
      const uint64_t cell = truncate_to_level(begin[cur].key, cand);
      size_t next = cur + 1;
      while (next < cnt && truncate_to_level(begin[next].key, cand) <= cell)
        ++next;
      const Box cb = get_bounds_from_morton_index(begin[cur].key, root, cand + 1);
      /* AABB::getCenter = min + extent()/2, math/AABB.h:70 */
      const double centre[3] = { cb.min[0] + (cb.max[0] - cb.min[0]) / 2,
                                 cb.min[1] + (cb.max[1] - cb.min[1]) / 2,
                                 cb.min[2] + (cb.max[2] - cb.min[2]) / 2 };
      size_t best = cur; /* std::min_element: first minimum */
      double best_d = sqdist(xyz + 3 * static_cast<uint64_t>(begin[cur].id), centre);
      for (size_t i = cur + 1; i < next; ++i) {
        const double d = sqdist(xyz + 3 * static_cast<uint64_t>(begin[i].id), centre);
        if (d < best_d) {
          best_d = d;
          best = i;
        }
      }
      sel[best] = 1;
      cur = next;
    }
    return stable_partition_flags(begin, end, sel);
  }

  /* PoissonDiskSampling::sample_points, Sampling.h:421-471 */
  size_t sample_min_distance(IP* begin,
                             IP* end,
                             uint64_t node_key,
                             int32_t node_level,
                             const Box& root,
                             float spacing_at_root)
  {
    const size_t cnt = static_cast<size_t>(end - begin);
    const Box nb = get_bounds_from_morton_index(node_key, root, static_cast<uint32_t>(node_level + 1));
    const double spacing_at_this_node = spacing_at_root / std::pow(2, node_level + 1);
    SparseGridRestated grid(nb, static_cast<float>(spacing_at_this_node));
    std::vector<uint8_t> sel(cnt, 0);
    for (size_t i = 0; i < cnt; ++i)
      sel[i] = grid.add(xyz + 3 * static_cast<uint64_t>(begin[i].id)) ? 1 : 0;
    return stable_partition_flags(begin, end, sel);
  }

  /* density_per_level of the MIN_DISTANCE_FAST strategy, process/TilerProcess.cpp:500-508 */
  static float density_per_level(int32_t node_level)
  {
    if (node_level < 0)
      return 0.25f;
    if (node_level < 1)
      return 0.5f;
    return 1.f;
  }

  /* AdaptivePoissonDiskSampling::sample_points, Sampling.h:477-542: only every n-th point of the
   * range is offered to the SparseGrid, the others are rejected outright */
  size_t sample_min_distance_fast(IP* begin,
                                  IP* end,
                                  uint64_t node_key,
                                  int32_t node_level,
                                  const Box& root,
                                  float spacing_at_root)
  {
    const size_t cnt = static_cast<size_t>(end - begin);
    const double spacing_at_this_node = spacing_at_root / std::pow(2, node_level + 1);
    const int cand = candidate_level_in_sampler(root, spacing_at_root, node_level);
    if (cand == -1)
      return 1; /* `return ++partition_point;` (Sampling.h:513-515) */
    const Box nb = get_bounds_from_morton_index(node_key, root, static_cast<uint32_t>(node_level + 1));
    SparseGridRestated grid(nb, static_cast<float>(spacing_at_this_node));
    const uint32_t nth_point = static_cast<uint32_t>(std::round(1 / density_per_level(node_level)));
    uint32_t point_counter = nth_point - 1; /* the first point is always analysed */
    std::vector<uint8_t> sel(cnt, 0);
    for (size_t i = 0; i < cnt; ++i) {
      if (++point_counter == nth_point) {
        point_counter = 0;
        sel[i] = grid.add(xyz + 3 * static_cast<uint64_t>(begin[i].id)) ? 1 : 0;
      }
    }
    return stable_partition_flags(begin, end, sel);
  }

  /* JitteredSampling::sample_points, Sampling.h:598-759 */
  size_t sample_jittered(IP* begin,
                         IP* end,
                         uint64_t node_key,
                         int32_t node_level,
                         const Box& root,
                         float spacing_at_root)
  {
    const size_t cnt = static_cast<size_t>(end - begin);
    const Box nb = get_bounds_from_morton_index(node_key, root, static_cast<uint32_t>(node_level + 1));
    const double spacing_at_this_node = spacing_at_root / std::pow(2, node_level + 1);
    const double perfect_cell_count = (nb.max[0] - nb.min[0]) / spacing_at_this_node;
    const uint32_t actual_cell_count = get_prev_power_of_two(static_cast<uint32_t>(perfect_cell_count));
    if (actual_cell_count < 16)
      throw OracleError(SW_ERR_JITTER_GRID_TOO_SMALL, "Grids smaller than 16x16 are not supported currently!");
    const uint32_t levels = static_cast<uint32_t>(std::log2(actual_cell_count));
    const uint32_t grid_level = static_cast<uint32_t>(node_level + levels);
    if (grid_level >= MAX_OCTREE_LEVELS)
      throw OracleError(SW_ERR_JITTER_NODE_TOO_SMALL, "Node is too small to be sampled with ImprovedPoissonSampling!");
    const uint64_t grid_mask = (1ull << (3 * levels)) - 1ull;
    const double grid_cell_size = (nb.max[0] - nb.min[0]) / actual_cell_count;
    const double permutation_cell_size = grid_cell_size / actual_cell_count;
    const uint32_t start_index = (3 * static_cast<uint32_t>(node_level + 1)) % 16;
    const uint32_t* perms[3];
    for (uint32_t a = 0; a < 3; ++a) {
      const uint32_t t = (a == 0) ? start_index : (start_index + a) % 16;
      if (actual_cell_count <= 16)
        perms[a] = PERMUTATIONS_16[t];
      else if (actual_cell_count <= 32)
        perms[a] = PERMUTATIONS_32[t];
      else
        perms[a] = PERMUTATIONS_64[t];
    }
    const uint32_t permutation_length = std::min<uint32_t>(actual_cell_count, 64);

    std::vector<uint8_t> sel(cnt, 0);
    size_t cur = 0;
    while (cur < cnt) {
      const uint64_t rel = truncate_to_level(begin[cur].key, grid_level);
      size_t next = cur + 1;
      while (next < cnt && truncate_to_level(begin[next].key, grid_level) <= rel)
        ++next;
      /* OctreeNodeIndex64::to_grid_index, OctreeNodeIndex.h:357-363 */
      const uint64_t idx = rel & grid_mask;
      const uint64_t lmask = (1u << levels) - 1;
      const uint64_t gz = contract_bits_by_3(idx) & lmask;
      const uint64_t gy = contract_bits_by_3(idx >> 1) & lmask;
      const uint64_t gx = contract_bits_by_3(idx >> 2) & lmask;
      /* uint32 table entry minus 1 stays uint32 (entries are >= 1) */
      const uint32_t px = perms[0][(gy + gz) % permutation_length] - 1;
      const uint32_t py = perms[1][(gx + gz) % permutation_length] - 1;
      const uint32_t pz = perms[2][(gx + gy) % permutation_length] - 1;
      const double target[3] = { nb.min[0] + (gx * grid_cell_size + px * permutation_cell_size),
                                 nb.min[1] + (gy * grid_cell_size + py * permutation_cell_size),
                                 nb.min[2] + (gz * grid_cell_size + pz * permutation_cell_size) };
      size_t best = cur;
      double best_d = sqdist(xyz + 3 * static_cast<uint64_t>(begin[cur].id), target);
      for (size_t i = cur + 1; i < next; ++i) {
        const double d = sqdist(xyz + 3 * static_cast<uint64_t>(begin[i].id), target);
        if (d < best_d) {
          best_d = d;
          best = i;
        }
      }
      sel[best] = 1;
      cur = next;
    }
    return stable_partition_flags(begin, end, sel);
  }
};

struct Handle
{
  sw_params params;
  std::vector<sw_node> nodes;
  std::vector<uint32_t> ids;
  std::vector<uint64_t> keys;
  std::vector<uint32_t> order;
  uint64_t duplicate_keys = 0;
  int32_t start_level = -1;
  double seconds = 0.0; /* index + sort + tiling only (o.run()), without building or copying buffers */
  std::string error;
};

} // namespace swo

using namespace swo;

extern "C" {

uint64_t
swo_expand_bits_by_3(uint64_t v)
{
  return expand_bits_by_3(v);
}

uint64_t
swo_contract_bits_by_3(uint64_t v)
{
  return contract_bits_by_3(v);
}

static Box
make_box(const double* bmin, const double* bmax)
{
  Box b;
  for (int a = 0; a < 3; ++a) {
    b.min[a] = bmin[a];
    b.max[a] = bmax[a];
  }
  return b;
}

/* index_point<21> over a batch; clamps xyz in place */
void
swo_index_points(double* xyz, uint64_t n, const double* bmin, const double* bmax, uint64_t* keys)
{
  const Box b = make_box(bmin, bmax);
  for (uint64_t i = 0; i < n; ++i)
    keys[i] = index_point(xyz + 3 * i, b);
}

void
swo_octant_bounds(uint8_t octant, const double* bmin, const double* bmax, double* out6)
{
  const Box r = get_octant_bounds(octant, make_box(bmin, bmax));
  std::memcpy(out6, r.min, 3 * sizeof(double));
  std::memcpy(out6 + 3, r.max, 3 * sizeof(double));
}

void
swo_bounds_from_morton_index(uint64_t key, uint32_t depth, const double* bmin, const double* bmax, double* out6)
{
  const Box r = get_bounds_from_morton_index(key, make_box(bmin, bmax), depth);
  std::memcpy(out6, r.min, 3 * sizeof(double));
  std::memcpy(out6 + 3, r.max, 3 * sizeof(double));
}

int32_t
swo_required_morton_index_depth(int32_t sampling,
                                int32_t node_level,
                                const double* bmin,
                                const double* bmax,
                                float root_max_spacing)
{
  RestatedPrims p{ nullptr, 0, sampling, 0 };
  NodeStructure root{};
  root.bounds = make_box(bmin, bmax);
  root.max_spacing = root_max_spacing;
  root.level = -1;
  return p.required_depth(node_level, root);
}

/* partition_points_into_child_octants over sorted keys: cuts[9] */
void
swo_partition_child_octants(const uint64_t* keys, uint64_t n, uint32_t level, uint64_t* cuts9)
{
  std::vector<IP> items(n);
  for (uint64_t i = 0; i < n; ++i)
    items[i] = { keys[i], static_cast<uint32_t>(i) };
  RestatedPrims p{ nullptr, n, 0, 0 };
  const auto c = p.partition(items.data(), items.data() + n, level);
  for (int i = 0; i < 9; ++i)
    cuts9[i] = c[i];
}

/*
 * sample_points() for one node (Sampling.h:799-821).  keys/ids: the node's points in Morton order
 * (ids index into xyz).  On return ids_out/keys_out hold [selected | unselected]; returns the
 * number selected, or -(error code).
 */
int64_t
swo_sample_points(int32_t sampling,
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
    std::vector<IP> items(n);
    for (uint64_t i = 0; i < n; ++i)
      items[i] = { keys[i], ids[i] };
    RestatedPrims p{ const_cast<double*>(xyz), 0, sampling, max_points_per_node };
    const size_t taken = p.sample(items.data(),
                                  items.data() + n,
                                  node_key,
                                  node_level,
                                  make_box(bmin, bmax),
                                  spacing_at_root,
                                  static_cast<Behaviour>(behaviour));
    for (uint64_t i = 0; i < n; ++i) {
      keys_out[i] = items[i].key;
      ids_out[i] = items[i].id;
    }
    return static_cast<int64_t>(taken);
  } catch (const OracleError& e) {
    return -static_cast<int64_t>(e.code);
  }
}

/* whole-batch tiling: ACCURATE (V1) or FAST (V3 first iteration + finalize) */
/* worker threads of the following swo_tile calls (1 = the plain sequential restatement) */
static unsigned g_threads = 1;

void
swo_set_threads(uint32_t n)
{
  g_threads = n ? n : 1;
}

/* 1 = sort with std::sort as the reference does (timing runs), 0 = std::stable_sort (parity runs, default) */
static int g_reference_sort = 0;

void
swo_set_reference_sort(int32_t on)
{
  g_reference_sort = on;
}

/* >= 0: FAST tiles with this start level instead of estimating it (subtree parity checks), -1 = estimate */
static int32_t g_start_level_override = -1;

void
swo_set_start_level_override(int32_t level)
{
  g_start_level_override = level;
}

/* `passes` runs of the whole path over the same points; seconds_out[k] = time of pass k, the handle holds the
 * result of the last pass (see swr_tile_repeat in ref_driver.cpp) */
int
swo_tile_repeat(const sw_params* params, double* xyz, uint64_t n, uint32_t passes, double* seconds_out,
                void** out_handle)
{
  auto* h = new Handle();
  h->params = *params;
  *out_handle = h;
  try {
    RestatedPrims prims{ xyz, n, params->sampling, params->max_points_per_node };
    for (uint32_t k = 0; k < (passes ? passes : 1u); ++k) {
      Orchestrator<RestatedPrims> o(prims, *params, g_threads);
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
    return SW_OK;
  } catch (const OracleError& e) {
    h->error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->error = e.what();
    return SW_ERR_STATE;
  }
}

int
swo_tile(const sw_params* params, double* xyz, uint64_t n, void** out_handle)
{
  return swo_tile_repeat(params, xyz, n, 1, nullptr, out_handle);
}

double
swo_tile_seconds(void* handle)
{
  return static_cast<Handle*>(handle)->seconds;
}

/* ACCURATE tiling of several batches, batch b = points [offsets[b], offsets[b + 1]) */
int
swo_tile_batches(const sw_params* params, double* xyz, uint64_t n, const uint64_t* offsets, uint32_t n_batches,
                 void** out_handle)
{
  auto* h = new Handle();
  h->params = *params;
  *out_handle = h;
  try {
    if (n_batches == 0 || offsets[n_batches] != n)
      throw OracleError(SW_ERR_INVALID_ARGUMENT, "tile_batches: offsets must end at n");
    RestatedPrims prims{ xyz, n, params->sampling, params->max_points_per_node };
    Orchestrator<RestatedPrims> o(prims, *params, 1);
    if (params->tiling == SW_FAST)
      o.run_fast_batches(offsets, n_batches);
    else
      o.run_accurate_batches(offsets, n_batches);
    h->start_level = o.start_level;
    h->nodes = std::move(o.nodes);
    h->ids = std::move(o.ids);
    h->keys = std::move(o.sorted_keys);
    h->order = std::move(o.sorted_ids);
    return SW_OK;
  } catch (const OracleError& e) {
    h->error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->error = e.what();
    return SW_ERR_STATE;
  }
}

uint64_t
swo_node_count(void* handle)
{
  return static_cast<Handle*>(handle)->nodes.size();
}

uint64_t
swo_point_id_count(void* handle)
{
  return static_cast<Handle*>(handle)->ids.size();
}

int32_t
swo_start_level(void* handle)
{
  return static_cast<Handle*>(handle)->start_level;
}

uint64_t
swo_duplicate_keys(void* handle)
{
  return static_cast<Handle*>(handle)->duplicate_keys;
}

void
swo_get_nodes(void* handle, sw_node* nodes, uint32_t* ids)
{
  auto* h = static_cast<Handle*>(handle);
  if (nodes)
    std::memcpy(nodes, h->nodes.data(), h->nodes.size() * sizeof(sw_node));
  if (ids)
    std::memcpy(ids, h->ids.data(), h->ids.size() * sizeof(uint32_t));
}

void
swo_get_keys(void* handle, uint64_t* keys, uint32_t* order)
{
  auto* h = static_cast<Handle*>(handle);
  if (keys)
    std::memcpy(keys, h->keys.data(), h->keys.size() * sizeof(uint64_t));
  if (order)
    std::memcpy(order, h->order.data(), h->order.size() * sizeof(uint32_t));
}

const char*
swo_last_error(void* handle)
{
  return static_cast<Handle*>(handle)->error.c_str();
}

void
swo_destroy(void* handle)
{
  delete static_cast<Handle*>(handle);
}

/* ---- LAS input transform and writer payloads ------------------------- */

/* position_from_las_point, core/io/LASFile.cpp:79-94, followed by the tiler's point transformation with
 * an identity SRS transform, core/process/TilerProcess.cpp:552-559 */
void
swo_las_positions(const int32_t* las, uint64_t n, const sw_las_transform* t, double* xyz)
{
  for (uint64_t i = 0; i < n; ++i) {
    for (int a = 0; a < 3; ++a) {
      double p = t->offset[a] + las[3 * i + a] * t->scale[a];
      p = std::min(t->header_max[a], std::max(t->header_min[a], p));
      if (t->shift_to_center) {
        p -= t->center[a];
        p = static_cast<float>(p);
      }
      xyz[3 * i + a] = p;
    }
  }
}

/* attributes::PositionAttribute::extractFromPoints(span<PointReference>), core/io/PNTSWriter.cpp:326-342:
 * the positions of the node's points, in the node's stored order, narrowed to float */
void
swo_payload_pnts(const double* xyz, const uint32_t* ids, uint64_t n_ids, float* out)
{
  for (uint64_t j = 0; j < n_ids; ++j)
    for (int a = 0; a < 3; ++a)
      out[3 * j + a] = static_cast<float>(xyz[3 * static_cast<uint64_t>(ids[j]) + a]);
}

/* compute_las_scale_from_bounds, core/io/LASPersistence.cpp:17-28 */
static double
las_scale_from_bounds(const Box& b)
{
  const double ex = b.max[0] - b.min[0], ey = b.max[1] - b.min[1], ez = b.max[2] - b.min[2];
  const double bounds_diagonal = std::sqrt(ex * ex + ey * ey + ez * ez);
  if (bounds_diagonal > 1000000)
    return 0.01;
  else if (bounds_diagonal > 100000)
    return 0.001;
  else if (bounds_diagonal > 1)
    return 0.001;
  return 0.0001;
}

/* LASzip (third-party, not vendored in /root/reference; the reference links the system liblaszip through
 * laszip_api.h): laszip_set_coordinates stores I32_QUANTIZE((coordinate - offset) / scale_factor) with
 * I32_QUANTIZE(n) = (n >= 0) ? (I32)(n + 0.5) : (I32)(n - 0.5)  (LASzip src/mydefs.hpp, laszip_dll.cpp). */
static int32_t
i32_quantize(double n)
{
  return (n >= 0) ? static_cast<int32_t>(n + 0.5) : static_cast<int32_t>(n - 0.5);
}

/* LASPersistence::persist_points, core/io/LASPersistence.h:119-131 (header: offset = min = bounds.min,
 * max = bounds.max, one scale) and :160-163 (laszip_set_coordinates of every point), for every node of a
 * node table; node bounds = get_bounds_from_node_index (OctreeAlgorithms.cpp:64-72). */
void
swo_payload_las(const double* xyz, const uint32_t* ids, const sw_node* nodes, uint64_t n_nodes, const double* bmin,
                const double* bmax, int32_t* out, sw_las_node_header* headers)
{
  const Box root = make_box(bmin, bmax);
  for (uint64_t r = 0; r < n_nodes; ++r) {
    Box b = root;
    for (uint32_t l = 0; l < nodes[r].levels; ++l)
      b = get_octant_bounds(static_cast<uint8_t>((nodes[r].index >> (3 * (nodes[r].levels - 1 - l))) & 7), b);
    const double scale = las_scale_from_bounds(b);
    if (headers) {
      for (int a = 0; a < 3; ++a) {
        headers[r].offset[a] = b.min[a];
        headers[r].max[a] = b.max[a];
      }
      headers[r].scale = scale;
      headers[r].reserved = 0;
    }
    for (uint64_t j = nodes[r].first; j < nodes[r].first + nodes[r].count; ++j)
      for (int a = 0; a < 3; ++a)
        out[3 * j + a] = i32_quantize((xyz[3 * static_cast<uint64_t>(ids[j]) + a] - b.min[a]) / scale);
  }
}

} /* extern "C" */
