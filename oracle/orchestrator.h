/*
 * oracle/orchestrator.h — TEST INFRASTRUCTURE ONLY (CPU oracle).  Nothing under oracle/ is on
 * the product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build or call it.
 *
 * CPU restatement of the reference's tiling orchestration for ONE batch
 * (internal_cache_size >= N, so no node is ever revisited).  Paths below are relative to
 * /root/reference/schwarzwald/core/tiling/.
 *
 *   tile_node              TilingAlgorithms.cpp:351-492   (terminal / internal / re-root branching)
 *   tile_internal_node     TilingAlgorithms.cpp:247-349
 *   tile_terminal_node     TilingAlgorithms.cpp:206-241
 *   split_range_into_child_nodes  TilingAlgorithms.cpp:116-162
 *   do_tiling_for_node     TilingAlgorithms.cpp:499-561   (task order is irrelevant to results)
 *   V1 (ACCURATE)          TilingAlgorithms.cpp:577-626
 *   V3 (FAST) first iteration     TilingAlgorithms.cpp:1250-1360
 *   estimate_start_node_level_in_octree   TilingAlgorithms.cpp:1473-1535
 *   split_indexed_points_into_subranges   TilingAlgorithms.cpp:1537-1578
 *   reconstruct_single_node / reconstruct_left_out_nodes  TilingAlgorithms.cpp:1661-1784
 *
 * The orchestration is a template over a `Prims` policy so that the same control flow can drive
 *   (a) the restated primitives in tiler_oracle.cpp (travels to the GPU box as source), and
 *   (b) the reference's own primitives compiled verbatim from /root/reference (ref_driver.cpp,
 *       built into oracle/_ref/, only buildable where /root/reference exists).
 * TilingAlgorithms.cpp itself cannot be compiled here (taskflow, boost::hana, cista are absent),
 * which is why the control flow is restated rather than linked.
 *
 * Sort tie rule (SURVEY.md §8a "S"): the reference uses unstable std::sort on the key only; the
 * oracle uses a stable sort, i.e. order by (key, original id).  The two agree whenever the batch
 * holds no duplicate 63-bit keys; `duplicate_keys` reports how many adjacent equal keys there were.
 *
 * Deviation: the deep re-root path (TilingAlgorithms.cpp:444-483) is not reproduced; hitting it
 * returns SW_ERR_DEEP_REROOT (the GPU library returns the same code).
 */
#pragma once

#include "../include/sw_types.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace swo {

struct OracleError : std::runtime_error
{
  int code;
  OracleError(int c, const std::string& msg)
    : std::runtime_error(msg)
    , code(c)
  {}
};

struct Box
{
  double min[3];
  double max[3];
};

/* octree::NodeStructure, Node.h:12-22 (name replaced by path index + levels) */
struct NodeStructure
{
  uint64_t morton_index; /* static 63-bit MortonIndex64 of the node */
  Box bounds;
  int32_t level; /* root = -1 */
  float max_spacing;
  uint32_t max_depth;
  uint64_t path_index; /* OctreeNodeIndex64::_index of the node's name */
  uint32_t path_levels;
};

constexpr uint32_t MAX_OCTREE_LEVELS = 21; /* TilingAlgorithms.cpp:20 */

enum Behaviour
{
  TakeAllWhenCountBelowMaxPoints = 0, /* Sampling.h:170-181 */
  AlwaysAdhereToMinSpacing = 1
};

/*
 * Prims policy requirements:
 *   using Item;                       // element being sorted (IndexedPoint64 analogue)
 *   uint64_t key(const Item&); uint32_t id(const Item&);
 *   void index_all(std::vector<Item>& out, const Box& bounds);   // index_point<21>, ClampToBounds
 *   void index_ids(const std::vector<uint32_t>& ids, std::vector<Item>& out, const Box& bounds);
 *   size_t sample(Item* begin, Item* end, uint64_t node_key, int32_t node_level,
 *                 const Box& root_bounds, float spacing_at_root, Behaviour b);  // sample_points
 *   int32_t required_depth(int32_t node_level, const NodeStructure& root);     // Sampling.cpp:29-62
 *   std::array<size_t, 9> partition(const Item* begin, const Item* end, uint32_t level);
 *   Box octant_bounds(uint8_t octant, const Box& parent);                       // OctreeAlgorithms.cpp:3-18
 */
template<class Prims>
struct Orchestrator
{
  using Item = typename Prims::Item;

  Prims& prims;
  sw_params params;
  std::vector<sw_node> nodes;
  std::vector<uint32_t> ids;
  std::vector<uint64_t> sorted_keys; /* test hook: keys after the sort */
  std::vector<uint32_t> sorted_ids;  /* test hook: sort permutation */
  uint64_t duplicate_keys = 0;
  int32_t start_level = -1; /* FAST: level_of_start_nodes */
  /* >= 0: FAST uses this start level instead of estimating it from the batch.  Checker hook for subtree parity
   * (bench.py, tests/test_gpu_fullsize.py): the points of a few level-3 subtrees of a large cloud are tiled
   * with the start level the WHOLE cloud produced, so that the nodes inside those subtrees are comparable. */
  int32_t start_level_override = -1;
  std::map<std::pair<uint32_t, uint64_t>, size_t> node_lookup;

  /* The reference runs the per-node work as taskflow tasks on its worker threads (one task per start node
   * in FAST, TilingAlgorithms.cpp:1314-1351; one subflow task per child node in ACCURATE, :499-561) and
   * indexes in parallel chunks (parallel::transform, :588-598); its sort is a single std::sort.  With
   * threads > 1 the same units of work run on std::threads here: every task writes into its own Collector,
   * collectors are merged in task order, so the result does not depend on the thread count. */
  unsigned threads = 1;

  struct Collector
  {
    std::vector<sw_node> nodes;
    std::vector<uint32_t> ids;
  };

  Orchestrator(Prims& p, const sw_params& prm, unsigned n_threads = 1)
    : prims(p)
    , params(prm)
    , threads(n_threads ? n_threads : 1)
  {}

  template<class Fn>
  void parallel_for(size_t n_tasks, Fn fn)
  {
    if (threads <= 1 || n_tasks <= 1) {
      for (size_t t = 0; t < n_tasks; ++t)
        fn(t);
      return;
    }
    std::atomic<size_t> next{ 0 };
    std::vector<std::exception_ptr> errors(threads);
    std::vector<std::thread> pool;
    const unsigned n_workers = static_cast<unsigned>(std::min<size_t>(threads, n_tasks));
    for (unsigned w = 0; w < n_workers; ++w)
      pool.emplace_back([&, w]() {
        try {
          for (size_t t = next.fetch_add(1); t < n_tasks; t = next.fetch_add(1))
            fn(t);
        } catch (...) {
          errors[w] = std::current_exception();
          next.store(n_tasks);
        }
      });
    for (auto& th : pool)
      th.join();
    for (auto& e : errors)
      if (e)
        std::rethrow_exception(e);
  }

  void merge(Collector& c)
  {
    const uint64_t base = ids.size();
    ids.insert(ids.end(), c.ids.begin(), c.ids.end());
    for (sw_node n : c.nodes) {
      n.first += base;
      node_lookup[{ n.levels, n.index }] = nodes.size();
      nodes.push_back(n);
    }
    c = Collector{};
  }

  Box root_bounds() const
  {
    Box b;
    for (int a = 0; a < 3; ++a) {
      b.min[a] = params.bounds_min[a];
      b.max[a] = params.bounds_max[a];
    }
    return b;
  }

  NodeStructure make_root() const
  {
    /* TilingAlgorithms.cpp:606-612 */
    NodeStructure r{};
    r.bounds = root_bounds();
    r.level = -1;
    r.max_depth = params.max_depth;
    r.max_spacing = params.spacing_at_root;
    r.morton_index = 0;
    r.path_index = 0;
    r.path_levels = 0;
    return r;
  }

  void persist(const Item* begin, const Item* end, const NodeStructure& node, uint32_t flags, Collector& out)
  {
    sw_node n{};
    n.index = node.path_index;
    n.levels = node.path_levels;
    n.flags = flags;
    n.first = out.ids.size();
    n.count = static_cast<uint64_t>(end - begin);
    for (const Item* it = begin; it != end; ++it)
      out.ids.push_back(prims.id(*it));
    out.nodes.push_back(n);
  }

  /* a subtree whose tiling was deferred to the worker threads (ACCURATE) */
  struct Deferred
  {
    Item* begin;
    Item* end;
    NodeStructure node;
  };

  /* tile_node + tile_internal_node + tile_terminal_node + recursion of do_tiling_for_node */
  void do_tiling_for_node(Item* begin, Item* end, const NodeStructure& node, const NodeStructure& root, Collector& out,
                          std::vector<Deferred>* defer = nullptr, int defer_below_level = 0)
  {
    const int32_t sample_level = prims.required_depth(node.level, root);
    const bool requires_deeper = sample_level > node.level;
    const int32_t max_level =
      static_cast<int32_t>(std::min<uint32_t>(MAX_OCTREE_LEVELS - 1, node.max_depth));

    bool terminal = false;
    if (!requires_deeper) {
      terminal = sample_level >= max_level; /* TilingAlgorithms.cpp:420-427 */
    } else {
      if (node.level >= max_level) { /* :436-442 */
        terminal = true;
      } else if (sample_level >= static_cast<int32_t>(MAX_OCTREE_LEVELS)) { /* :444-483 */
        throw OracleError(SW_ERR_DEEP_REROOT, "deep re-root path is not supported");
      }
    }

    if (terminal) {
      persist(begin, end, node, SW_NODE_TERMINAL, out);
      return;
    }

    if (begin == end) /* :253-259 */
      throw OracleError(SW_ERR_EMPTY_NODE, "tile_internal_node: Got zero points to tile");

    /* no cached points in a single-batch run -> TakeAllWhenCountBelowMaxPoints (:272-275) */
    const int32_t rel_level = node.level - (root.level + 1);
    const size_t taken = prims.sample(begin,
                                      end,
                                      node.morton_index,
                                      rel_level,
                                      root.bounds,
                                      root.max_spacing,
                                      TakeAllWhenCountBelowMaxPoints);
    const bool took_all = (begin + taken == end);
    const bool by_count = static_cast<uint64_t>(end - begin) <= params.max_points_per_node;
    persist(begin, begin + taken, node, (took_all && by_count) ? SW_NODE_TAKE_ALL : 0u, out);

    /* split_range_into_child_nodes, :116-162 */
    Item* rest = begin + taken;
    if (rest == end)
      return;
    const int32_t child_level = node.level + 1;
    if (child_level >= static_cast<int32_t>(MAX_OCTREE_LEVELS))
      throw OracleError(SW_ERR_DEEP_REROOT, "child level exceeds MortonIndex64 capacity");
    const auto cuts = prims.partition(rest, end, static_cast<uint32_t>(child_level));
    for (uint8_t octant = 0; octant < 8; ++octant) {
      if (cuts[octant] == cuts[octant + 1])
        continue;
      NodeStructure child = node;
      const uint32_t shift = (MAX_OCTREE_LEVELS - child_level - 1) * 3;
      child.morton_index |= (static_cast<uint64_t>(octant & 7) << shift);
      child.bounds = prims.octant_bounds(octant, node.bounds);
      child.level = child_level;
      child.max_spacing /= 2;
      child.path_index = (node.path_index << 3) | octant;
      child.path_levels = node.path_levels + 1;
      if (defer && child.level >= defer_below_level)
        defer->push_back({ rest + cuts[octant], rest + cuts[octant + 1], child });
      else
        do_tiling_for_node(rest + cuts[octant], rest + cuts[octant + 1], child, root, out, defer, defer_below_level);
    }
  }

  /* The reference calls std::sort (TilingAlgorithms.cpp:600-604, 1289-1292), whose order of equal keys is
   * unspecified; parity runs pin it to "stable by original index"  with std::stable_sort.
   * Timing runs of the reference arm use std::sort itself. */
  bool reference_sort = false;

  void sort_items(std::vector<Item>& items)
  {
    auto by_key = [this](const Item& l, const Item& r) { return prims.key(l) < prims.key(r); };
    if (reference_sort)
      std::sort(items.begin(), items.end(), by_key);
    else
      std::stable_sort(items.begin(), items.end(), by_key);
    sorted_keys.resize(items.size());
    sorted_ids.resize(items.size());
    duplicate_keys = 0;
    for (size_t i = 0; i < items.size(); ++i) {
      sorted_keys[i] = prims.key(items[i]);
      sorted_ids[i] = prims.id(items[i]);
      if (i && sorted_keys[i] == sorted_keys[i - 1])
        ++duplicate_keys;
    }
  }

  /* estimate_start_node_level_in_octree, :1473-1535, on the sorted keys */
  uint32_t estimate_start_level(const std::vector<Item>& items, size_t concurrency)
  {
    using Range = std::pair<size_t, size_t>;
    std::vector<Range> splits{ { 0, items.size() } };
    constexpr uint32_t MIN_LEVEL = 3, MAX_LEVEL = 6;
    for (uint32_t level = 0; level < MAX_LEVEL; ++level) {
      std::vector<Range> next;
      for (const auto& r : splits) {
        const auto cuts = prims.partition(items.data() + r.first, items.data() + r.second, level);
        for (int o = 0; o < 8; ++o)
          if (cuts[o + 1] > cuts[o])
            next.push_back({ r.first + cuts[o], r.first + cuts[o + 1] });
      }
      splits.swap(next);
      float score = 0.f;
      if (!(splits.size() <= concurrency / 2)) {
        size_t large = 0;
        for (const auto& r : splits)
          if (r.second - r.first >= 100000)
            ++large;
        score = static_cast<float>(large) / static_cast<float>(concurrency);
      }
      if (score >= 1.f)
        return std::max(level + 1, MIN_LEVEL);
    }
    return MAX_LEVEL;
  }

  /* parallel::transform over chunks of the batch (TilingAlgorithms.cpp:588-598, :1262-1285) */
  void index_items(std::vector<Item>& items)
  {
    if (threads <= 1) {
      prims.index_all(items, root_bounds());
      return;
    }
    const size_t n = prims.point_count();
    const size_t n_chunks = std::min<size_t>(threads, std::max<size_t>(1, n / 65536));
    std::vector<std::vector<Item>> parts(n_chunks);
    const Box rb = root_bounds();
    parallel_for(n_chunks, [&](size_t c) {
      prims.index_range(n * c / n_chunks, n * (c + 1) / n_chunks, parts[c], rb);
    });
    items.clear();
    items.reserve(n);
    for (auto& part : parts)
      items.insert(items.end(), part.begin(), part.end());
  }

  void run_accurate()
  {
    std::vector<Item> items;
    index_items(items);
    sort_items(items);
    if (items.empty()) /* the reference would throw in tile_internal_node */
      throw OracleError(SW_ERR_EMPTY_NODE, "tile_internal_node: Got zero points to tile");
    const NodeStructure root = make_root();
    Collector top;
    if (threads <= 1) {
      do_tiling_for_node(items.data(), items.data() + items.size(), root, root, top);
      merge(top);
      return;
    }
    /* the root and its children are single tasks in the reference as well; their subtrees (nodes of
     * level >= 1) become independent tasks */
    std::vector<Deferred> tasks;
    do_tiling_for_node(items.data(), items.data() + items.size(), root, root, top, &tasks, 1);
    merge(top);
    std::vector<Collector> outs(tasks.size());
    parallel_for(tasks.size(), [&](size_t t) {
      do_tiling_for_node(tasks[t].begin, tasks[t].end, tasks[t].node, root, outs[t]);
    });
    for (auto& c : outs)
      merge(c);
  }

  void run_fast()
  {
    std::vector<Item> items;
    index_items(items);
    /* parallel::scatter throws when there are fewer points than tasks (threading/Parallel.h:181-186) */
    if (items.size() < params.concurrency)
      throw OracleError(SW_ERR_TOO_FEW_POINTS, "fewer points than indexing threads");
    sort_items(items);
    const uint32_t S = start_level_override >= 0 ? static_cast<uint32_t>(start_level_override)
                                                 : estimate_start_level(items, params.concurrency);
    start_level = static_cast<int32_t>(S);
    const NodeStructure root = make_root();

    /* split_indexed_points_into_subranges + one task per non-empty level-S node, :1302-1351 */
    const uint32_t shift = (MAX_OCTREE_LEVELS - S) * 3;
    size_t b = 0;
    std::vector<uint64_t> start_nodes;
    std::vector<Deferred> tasks;
    while (b < items.size()) {
      const uint64_t prefix = prims.key(items[b]) >> shift;
      size_t e = b + 1;
      while (e < items.size() && (prims.key(items[e]) >> shift) == prefix)
        ++e;
      NodeStructure n{};
      n.bounds = root.bounds;
      for (uint32_t l = 0; l < S; ++l) /* get_bounds_from_node_index, OctreeAlgorithms.cpp:64-72 */
        n.bounds = prims.octant_bounds(static_cast<uint8_t>((prefix >> (3 * (S - 1 - l))) & 7), n.bounds);
      n.level = static_cast<int32_t>(S) - 1;
      n.max_depth = root.max_depth;
      n.max_spacing = static_cast<float>(root.max_spacing / std::pow(2, S));
      n.morton_index = prefix << shift;
      n.path_index = prefix;
      n.path_levels = S;
      start_nodes.push_back(prefix);
      tasks.push_back({ items.data() + b, items.data() + e, n });
      b = e;
    }
    std::vector<Collector> outs(tasks.size());
    parallel_for(tasks.size(), [&](size_t t) {
      do_tiling_for_node(tasks[t].begin, tasks[t].end, tasks[t].node, root, outs[t]);
    });
    for (auto& c : outs)
      merge(c);
    reconstruct(S, start_nodes);
  }

  /* reconstruct_left_out_nodes + reconstruct_single_node, :1661-1784 */
  void reconstruct(uint32_t S, const std::vector<uint64_t>& start_nodes)
  {
    if (S == 0)
      return;
    const Box rb = root_bounds();
    for (int32_t lv = static_cast<int32_t>(S) - 1; lv >= 0; --lv) {
      /* ancestors with `lv` levels of every existing start node, deepest level first */
      std::vector<uint64_t> parents;
      for (uint64_t p : start_nodes)
        parents.push_back(p >> (3 * (S - lv)));
      std::sort(parents.begin(), parents.end());
      parents.erase(std::unique(parents.begin(), parents.end()), parents.end());
      std::vector<Collector> outs(parents.size());
      parallel_for(parents.size(), [&](size_t pi) {
        const uint64_t parent = parents[pi];
        std::vector<uint32_t> child_ids;
        for (uint8_t octant = 0; octant < 8; ++octant) {
          auto it = node_lookup.find({ static_cast<uint32_t>(lv + 1), (parent << 3) | octant });
          if (it == node_lookup.end())
            continue;
          const sw_node& cn = nodes[it->second];
          child_ids.insert(child_ids.end(), ids.begin() + cn.first, ids.begin() + cn.first + cn.count);
        }
        std::vector<Item> items;
        prims.index_ids(child_ids, items, rb);
        /* MemoryPersistence::is_lossless() == true -> no re-sort (:1689-1691) */
        const uint32_t shift = (MAX_OCTREE_LEVELS - lv) * 3;
        const uint64_t node_key = (lv == 0) ? 0 : (parent << shift);
        const size_t taken = prims.sample(items.data(),
                                          items.data() + items.size(),
                                          node_key,
                                          lv - 1,
                                          rb,
                                          params.spacing_at_root,
                                          AlwaysAdhereToMinSpacing);
        NodeStructure n{};
        n.path_index = parent;
        n.path_levels = static_cast<uint32_t>(lv);
        persist(items.data(), items.data() + taken, n, SW_NODE_RECONSTRUCTED, outs[pi]);
      });
      for (auto& c : outs)
        merge(c);
    }
  }

  /* ---- several batches through TilingAlgorithmV1 (ACCURATE) ---------------------------
   * The reference keeps what every node stores in its persistence and, when a later batch reaches the node
   * again, reads those points back, merges them with the incoming ones and samples the union
   * (tile_node, TilingAlgorithms.cpp:351-492).  Restated here with a lossless in-memory store (MemoryPersistence:
   * is_lossless() == true, so read_pnts_from_disk does not re-sort, :104-106).
   * Extra Prims requirements:
   *   uint64_t morton_in_bounds(uint32_t id, const Box& bounds);   // calculate_morton_index<21>, no clamping
   *   Item make_item(uint32_t id, uint64_t key);
   */
  std::map<std::pair<uint32_t, uint64_t>, std::vector<uint32_t>> store; /* node -> stored ids, stored order */
  std::map<std::pair<uint32_t, uint64_t>, uint32_t> store_flags;

  void store_node(const Item* begin, const Item* end, const NodeStructure& node, uint32_t flags)
  {
    std::vector<uint32_t>& v = store[{ node.path_levels, node.path_index }];
    v.clear();
    for (const Item* it = begin; it != end; ++it)
      v.push_back(prims.id(*it));
    store_flags[{ node.path_levels, node.path_index }] = flags;
  }

  void tile_node_cached(std::vector<Item>& node_data, const NodeStructure& node, const NodeStructure& root)
  {
    /* read_pnts_from_disk, :50-109: the node's key is the baseline, only the levels below the node are
     * recomputed, relative to the NODE's bounds */
    std::vector<Item> cached;
    const auto found = store.find({ node.path_levels, node.path_index });
    if (found != store.end()) {
      const uint32_t start_level = static_cast<uint32_t>(node.level + 1);
      cached.reserve(found->second.size());
      for (uint32_t id : found->second) {
        const uint64_t below = prims.morton_in_bounds(id, node.bounds);
        /* set_octant_at_level(level, below.get_octant_at_level(level - start_level)) for level >= start_level:
         * the top (21 - start_level) levels of `below` move down behind the node's own levels */
        const uint64_t key = node.morton_index | (start_level < MAX_OCTREE_LEVELS ? (below >> (3 * start_level)) : 0);
        cached.push_back(prims.make_item(id, key));
      }
    }
    const size_t cached_count = cached.size();

    const int32_t sample_level = prims.required_depth(node.level, root);
    const bool requires_deeper = sample_level > node.level;
    const int32_t max_level = static_cast<int32_t>(std::min<uint32_t>(MAX_OCTREE_LEVELS - 1, node.max_depth));
    bool terminal = false;
    if (!requires_deeper) {
      terminal = sample_level >= max_level; /* :420-427 */
    } else {
      if (node.level >= max_level) /* :436-442 */
        terminal = true;
      else if (sample_level >= static_cast<int32_t>(MAX_OCTREE_LEVELS)) /* :444-483 */
        throw OracleError(SW_ERR_DEEP_REROOT, "deep re-root path is not supported");
    }

    std::vector<Item> all;
    if (terminal) { /* merge_node_data_unsorted, Node.cpp:22-34 */
      if (node_data.empty())
        all = std::move(cached);
      else {
        all = std::move(node_data);
        all.insert(all.end(), cached.begin(), cached.end());
      }
      store_node(all.data(), all.data() + all.size(), node, SW_NODE_TERMINAL);
      return;
    }
    /* merge_node_data_sorted, Node.cpp:3-20: incoming points first on equal keys */
    if (node_data.empty())
      all = std::move(cached);
    else if (cached.empty())
      all = std::move(node_data);
    else {
      all.reserve(node_data.size() + cached.size());
      std::merge(node_data.begin(), node_data.end(), cached.begin(), cached.end(), std::back_inserter(all),
                 [this](const Item& l, const Item& r) { return prims.key(l) < prims.key(r); });
    }
    if (all.empty()) /* :253-259 */
      throw OracleError(SW_ERR_EMPTY_NODE, "tile_internal_node: Got zero points to tile");

    /* once a node has been sampled it is always sampled again (:272-275) */
    const Behaviour behaviour = cached_count > 0 ? AlwaysAdhereToMinSpacing : TakeAllWhenCountBelowMaxPoints;
    const int32_t rel_level = node.level - (root.level + 1);
    const size_t taken = prims.sample(all.data(), all.data() + all.size(), node.morton_index, rel_level, root.bounds,
                                      root.max_spacing, behaviour);
    const bool took_all = taken == all.size();
    const bool by_count = behaviour == TakeAllWhenCountBelowMaxPoints && all.size() <= params.max_points_per_node;
    store_node(all.data(), all.data() + taken, node, (took_all && by_count) ? SW_NODE_TAKE_ALL : 0u);

    /* split_range_into_child_nodes, :116-162: everything that was not selected moves down, points this node
     * stored in an earlier batch included */
    Item* rest = all.data() + taken;
    Item* end = all.data() + all.size();
    if (rest == end)
      return;
    const int32_t child_level = node.level + 1;
    if (child_level >= static_cast<int32_t>(MAX_OCTREE_LEVELS))
      throw OracleError(SW_ERR_DEEP_REROOT, "child level exceeds MortonIndex64 capacity");
    const auto cuts = prims.partition(rest, end, static_cast<uint32_t>(child_level));
    for (uint8_t octant = 0; octant < 8; ++octant) {
      if (cuts[octant] == cuts[octant + 1])
        continue;
      NodeStructure child = node;
      const uint32_t shift = (MAX_OCTREE_LEVELS - child_level - 1) * 3;
      child.morton_index |= (static_cast<uint64_t>(octant & 7) << shift);
      child.bounds = prims.octant_bounds(octant, node.bounds);
      child.level = child_level;
      child.max_spacing /= 2;
      child.path_index = (node.path_index << 3) | octant;
      child.path_levels = node.path_levels + 1;
      std::vector<Item> child_data(rest + cuts[octant], rest + cuts[octant + 1]);
      tile_node_cached(child_data, child, root);
    }
  }

  /* batch b = points [offsets[b], offsets[b + 1]) of the prims' point array; one build_execution_graph per batch
   * (TilingAlgorithms.cpp:577-626) */
  void run_accurate_batches(const uint64_t* offsets, uint32_t n_batches)
  {
    const NodeStructure root = make_root();
    const Box rb = root_bounds();
    for (uint32_t b = 0; b < n_batches; ++b) {
      std::vector<Item> items;
      prims.index_range(offsets[b], offsets[b + 1], items, rb);
      sort_items(items);
      if (items.empty())
        throw OracleError(SW_ERR_EMPTY_NODE, "tile_internal_node: Got zero points to tile");
      tile_node_cached(items, root, root);
    }
    emit_store();
  }

  void emit_store()
  {
    for (const auto& kv : store) { /* the final content of the persistence, (levels, index) order */
      sw_node n{};
      n.levels = kv.first.first;
      n.index = kv.first.second;
      n.flags = store_flags[kv.first];
      n.first = ids.size();
      n.count = kv.second.size();
      ids.insert(ids.end(), kv.second.begin(), kv.second.end());
      node_lookup[kv.first] = nodes.size();
      nodes.push_back(n);
    }
  }

  /* TilingAlgorithmV3 over several batches: the first batch fixes the start level
   * (build_execution_graph_for_first_iteration, :1250-1360); later batches are indexed and sorted in
   * num_indexing_threads chunks, split at that level and k-way merged per start node
   * (build_execution_graph_for_later_iterations, :1362-1453; merge_ranges, Algorithm.h:111-150, keeps the
   * earlier chunk on equal keys) -- with the tie rule "stable by original index" that is the stable sort of the
   * whole batch cut at the start nodes.  Every start node goes through tile_node with its cached points;
   * finalize reconstructs the ancestors of every start node the persistence holds (:1717-1784). */
  void run_fast_batches(const uint64_t* offsets, uint32_t n_batches)
  {
    const NodeStructure root = make_root();
    const Box rb = root_bounds();
    uint32_t S = 0;
    for (uint32_t b = 0; b < n_batches; ++b) {
      std::vector<Item> items;
      prims.index_range(offsets[b], offsets[b + 1], items, rb);
      if (items.size() < params.concurrency) /* parallel::scatter, threading/Parallel.h:181-186 */
        throw OracleError(SW_ERR_TOO_FEW_POINTS, "fewer points than indexing threads");
      sort_items(items);
      if (b == 0) {
        S = estimate_start_level(items, params.concurrency);
        start_level = static_cast<int32_t>(S);
      }
      const uint32_t shift = (MAX_OCTREE_LEVELS - S) * 3;
      size_t lo = 0;
      while (lo < items.size()) {
        const uint64_t prefix = prims.key(items[lo]) >> shift;
        size_t hi = lo + 1;
        while (hi < items.size() && (prims.key(items[hi]) >> shift) == prefix)
          ++hi;
        NodeStructure n{}; /* prepare_range_for_tiling, :1622-1660 */
        n.bounds = root.bounds;
        for (uint32_t l = 0; l < S; ++l)
          n.bounds = prims.octant_bounds(static_cast<uint8_t>((prefix >> (3 * (S - 1 - l))) & 7), n.bounds);
        n.level = static_cast<int32_t>(S) - 1;
        n.max_depth = root.max_depth;
        n.max_spacing = static_cast<float>(root.max_spacing / std::pow(2, S));
        n.morton_index = prefix << shift;
        n.path_index = prefix;
        n.path_levels = S;
        std::vector<Item> node_items(items.begin() + lo, items.begin() + hi);
        tile_node_cached(node_items, n, root);
        lo = hi;
      }
    }
    /* finalize: reconstruct_left_out_nodes over every start node that exists */
    for (int32_t lv = static_cast<int32_t>(S) - 1; lv >= 0; --lv) {
      std::vector<uint64_t> parents;
      for (const auto& kv : store)
        if (kv.first.first == S)
          parents.push_back(kv.first.second >> (3 * (S - lv)));
      std::sort(parents.begin(), parents.end());
      parents.erase(std::unique(parents.begin(), parents.end()), parents.end());
      for (uint64_t parent : parents) {
        std::vector<uint32_t> child_ids;
        for (uint8_t octant = 0; octant < 8; ++octant) {
          const auto it = store.find({ static_cast<uint32_t>(lv + 1), (parent << 3) | octant });
          if (it != store.end())
            child_ids.insert(child_ids.end(), it->second.begin(), it->second.end());
        }
        std::vector<Item> items;
        prims.index_ids(child_ids, items, rb);
        const uint32_t shift = (MAX_OCTREE_LEVELS - lv) * 3;
        const uint64_t node_key = (lv == 0) ? 0 : (parent << shift);
        const size_t taken = prims.sample(items.data(), items.data() + items.size(), node_key, lv - 1, rb,
                                          params.spacing_at_root, AlwaysAdhereToMinSpacing);
        NodeStructure n{};
        n.path_index = parent;
        n.path_levels = static_cast<uint32_t>(lv);
        store_node(items.data(), items.data() + taken, n, SW_NODE_RECONSTRUCTED);
      }
    }
    emit_store();
  }

  void run()
  {
    if (params.tiling == SW_ACCURATE)
      run_accurate();
    else if (params.tiling == SW_FAST)
      run_fast();
    else
      throw OracleError(SW_ERR_INVALID_ARGUMENT, "unknown tiling strategy");
  }
};

} // namespace swo
