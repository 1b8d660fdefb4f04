// kernels_sampling.cu — the level-synchronous sampling sweep (K3, K5, K6, K7, K8).
//
// The reference walks the octree node by node (do_tiling_for_node, tiling/TilingAlgorithms.cpp:
// 499-561): sample a node, persist the selected points, split the rest into the eight children.
// Node outputs are independent of each other, and what a node does depends only on its level and
// its point count.  So the GPU processes one octree LEVEL at a time over the whole Morton-ordered
// remainder list:
//
//   node_rle        K3  node boundaries = runs of equal (key >> node_shift)        [read 8 B/pt]
//                       (partition_points_into_child_octants, OctreeAlgorithms.h:240-265)
//   select_*        K6/K7 per-cell selection (RANDOM_GRID is folded into the compaction)
//   level_compact   K5+K8 take-all decision (Sampling.h:201-208), stable partition
//                       [selected | remainder] (stable_partition_with_jumps, util/algorithms/
//                       Algorithm.h:22-77), node table rows                          [read 12, write 12 B/pt]
//
// The remainder list of level L is the input list of level L+1.  All kernels use the same tile
// geometry (SW_SWEEP_TILE elements, warp-striped, 8 items per lane) and hand tiles out through an
// atomic ticket so that the decoupled look-back chains are dead-lock free.
#include "swgpu_internal.cuh"

#define SWP_THREADS 256
#define SWP_WARPS (SWP_THREADS / 32)
#define SWP_ITEMS 8
static_assert(SWP_THREADS * SWP_ITEMS == SW_SWEEP_TILE, "tile geometry");

size_t
sweep_tiles(u64 count)
{
  const size_t t = (size_t)((count + SW_SWEEP_TILE - 1) / SW_SWEEP_TILE);
  return t ? t : 1;
}

// position of item j of this lane inside the tile (warp-striped)
__device__ __forceinline__ u32
item_pos(u32 warp, u32 lane, int j)
{
  return warp * (32 * SWP_ITEMS) + j * 32 + lane;
}

__device__ __forceinline__ u32
take_ticket(u32* ticket, u32* s_slot)
{
  if (threadIdx.x == 0)
    *s_slot = atomicAdd(ticket, 1u);
  __syncthreads();
  return *s_slot;
}

// exclusive prefix of one u32 per warp over the warps of the block + block total.
// s_w must hold SWP_WARPS entries; contains a __syncthreads().
__device__ __forceinline__ u32
warp_totals_exclusive(u32 my_warp_total, u32 warp, u32 lane, u32* s_w, u32& block_total)
{
  if (lane == 0)
    s_w[warp] = my_warp_total;
  __syncthreads();
  u32 excl = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SWP_WARPS; ++w) {
    const u32 v = s_w[w];
    excl += (w < (int)warp) ? v : 0u;
    tot += v;
  }
  block_total = tot;
  return excl;
}

// =============================================================================================
// K3  node run-length encoder
// =============================================================================================
__global__ void __launch_bounds__(SWP_THREADS)
node_rle_kernel(const u64* __restrict__ keys, u64 count, int node_shift, u32* __restrict__ node_start,
                u32* __restrict__ tile_rank0, u32* __restrict__ n_nodes, u64* __restrict__ status,
                u32* __restrict__ ticket)
{
  __shared__ u32 s_slot;
  __shared__ u32 s_w[SWP_WARPS];
  __shared__ u64 s_prefix;
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 tile = take_ticket(ticket, &s_slot);
  const u64 base = (u64)tile * SW_SWEEP_TILE;

  u32 hmask[SWP_ITEMS];
  u32 wcount = 0;
#pragma unroll
  for (int j = 0; j < SWP_ITEMS; ++j) {
    const u64 i = base + item_pos(warp, lane, j);
    bool head = false;
    if (i < count) {
      const u64 k = keys[i] & SW_KEY_MASK;
      head = (i == 0) || ((k >> node_shift) != ((keys[i - 1] & SW_KEY_MASK) >> node_shift));
    }
    hmask[j] = __ballot_sync(0xffffffffu, head);
    wcount += __popc(hmask[j]);
  }
  u32 total;
  const u32 wexcl = warp_totals_exclusive(wcount, warp, lane, s_w, total);
  if (warp == 0) {
    const u64 p = lookback_exclusive(status, tile, (u64)total);
    if (lane == 0)
      s_prefix = p;
  }
  __syncthreads();
  const u32 prefix = (u32)s_prefix;
  if (threadIdx.x == 0) {
    tile_rank0[tile] = prefix;
    if (base + SW_SWEEP_TILE >= count) { // last tile
      *n_nodes = prefix + total;
      node_start[prefix + total] = (u32)count; // sentinel
    }
  }
  u32 run = prefix + wexcl;
  const u32 lt = lanemask_lt();
#pragma unroll
  for (int j = 0; j < SWP_ITEMS; ++j) {
    if ((hmask[j] >> lane) & 1u)
      node_start[run + __popc(hmask[j] & lt)] = (u32)(base + item_pos(warp, lane, j));
    run += __popc(hmask[j]);
  }
}

void
launch_node_rle(const u64* keys, u64 count, int node_shift, u32* node_start, u32* tile_rank0, u32* n_nodes,
                u64* status, u32* ticket, cudaStream_t stream)
{
  const size_t tiles = sweep_tiles(count);
  cudaMemsetAsync(status, 0, tiles * sizeof(u64), stream);
  cudaMemsetAsync(ticket, 0, sizeof(u32), stream);
  node_rle_kernel<<<(u32)tiles, SWP_THREADS, 0, stream>>>(keys, count, node_shift, node_start, tile_rank0, n_nodes,
                                                         status, ticket);
}

// =============================================================================================
// K5 + K6 + K8  take-all decision, RANDOM_GRID selection, stable two-way compaction
// =============================================================================================
__global__ void __launch_bounds__(SWP_THREADS)
level_compact_kernel(SwLevelArgs a, u64* __restrict__ n_selected, u64* __restrict__ status, u32* __restrict__ ticket)
{
  __shared__ u32 s_slot;
  __shared__ u32 s_w[SWP_WARPS];
  __shared__ u32 s_w2[SWP_WARPS];
  __shared__ u64 s_prefix;
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 tile = take_ticket(ticket, &s_slot);
  const u64 base = (u64)tile * SW_SWEEP_TILE;
  const u32 lt = lanemask_lt();

  // ---- phase 1: keys, node heads, cell heads ---------------------------------------------------
  u64 key[SWP_ITEMS];
  u32 nmask[SWP_ITEMS];
  u32 cmask[SWP_ITEMS];
  u32 wheads = 0;
#pragma unroll
  for (int j = 0; j < SWP_ITEMS; ++j) {
    const u64 i = base + item_pos(warp, lane, j);
    bool nh = false, ch = false;
    key[j] = 0;
    if (i < a.count) {
      const u64 k = a.in_key[i] & SW_KEY_MASK;
      key[j] = k;
      if (i == 0) {
        nh = ch = true;
      } else {
        const u64 pk = a.in_key[i - 1] & SW_KEY_MASK;
        nh = (k >> a.node_shift) != (pk >> a.node_shift);
        ch = nh || ((k >> a.cell_shift) != (pk >> a.cell_shift));
      }
    }
    nmask[j] = __ballot_sync(0xffffffffu, nh);
    cmask[j] = __ballot_sync(0xffffffffu, ch);
    wheads += __popc(nmask[j]);
  }
  u32 heads_total;
  const u32 hexcl = warp_totals_exclusive(wheads, warp, lane, s_w, heads_total);

  // ---- phase 2: node rank -> take-all decision -> selection flags ------------------------------
  u32 node_rank[SWP_ITEMS];
  u32 smask[SWP_ITEMS];
  u32 wsel = 0;
  {
    u32 run = a.tile_rank0[tile] + hexcl; // heads before this item
#pragma unroll
    for (int j = 0; j < SWP_ITEMS; ++j) {
      const u64 i = base + item_pos(warp, lane, j);
      const u32 incl = run + __popc(nmask[j] & (lt | (1u << lane)));
      node_rank[j] = incl - 1;
      run += __popc(nmask[j]);
      bool sel = false;
      if (i < a.count) {
        bool take = a.force_all != 0;
        if (!take && a.allow_take_all) {
          const u32 cnt = a.node_start[node_rank[j] + 1] - a.node_start[node_rank[j]];
          take = (u64)cnt <= a.max_points_per_node;
        }
        if (take)
          sel = true;
        else if (a.sampling == SW_RANDOM_GRID)
          sel = (cmask[j] >> lane) & 1u; // first point of each cell run (Sampling.h:253-284)
        else
          sel = a.sel[i] != 0;
      }
      smask[j] = __ballot_sync(0xffffffffu, sel);
      wsel += __popc(smask[j]);
    }
  }
  u32 sel_total;
  const u32 sexcl = warp_totals_exclusive(wsel, warp, lane, s_w2, sel_total);
  if (warp == 0) {
    const u64 p = lookback_exclusive(status, tile, (u64)sel_total);
    if (lane == 0)
      s_prefix = p;
  }
  __syncthreads();
  const u64 sel_prefix = s_prefix;
  if (threadIdx.x == 0 && base + SW_SWEEP_TILE >= a.count)
    *n_selected = sel_prefix + sel_total;

  // ---- phase 3: scatter ------------------------------------------------------------------------
  u64 run = sel_prefix + sexcl; // selected before this item
#pragma unroll
  for (int j = 0; j < SWP_ITEMS; ++j) {
    const u64 i = base + item_pos(warp, lane, j);
    if (i < a.count) {
      const u64 srank = run + __popc(smask[j] & lt);
      const u32 idx = a.in_idx ? a.in_idx[i] : (u32)i;
      if ((smask[j] >> lane) & 1u) {
        a.out_key[a.out_offset + srank] = key[j];
        a.out_idx[a.out_offset + srank] = idx;
      } else if (a.rem_key) {
        a.rem_key[i - srank] = key[j];
        a.rem_idx[i - srank] = idx;
      }
      if ((nmask[j] >> lane) & 1u) {
        a.node_index[a.node_base + node_rank[j]] = key[j] >> a.node_shift;
        a.node_first[a.node_base + node_rank[j]] = a.out_offset + srank;
      }
    }
    run += __popc(smask[j]);
  }
}

void
launch_level_compact(const SwLevelArgs& a, u64* n_selected, u64* status, u32* ticket, cudaStream_t stream)
{
  const size_t tiles = sweep_tiles(a.count);
  cudaMemsetAsync(status, 0, tiles * sizeof(u64), stream);
  cudaMemsetAsync(ticket, 0, sizeof(u32), stream);
  level_compact_kernel<<<(u32)tiles, SWP_THREADS, 0, stream>>>(a, n_selected, status, ticket);
}

// =============================================================================================
// K7  GRID_CENTER / JITTERED: segmented first-arg-min of the squared distance to a per-cell target
// =============================================================================================
//   GridCenterSampling::sample_points   tiling/Sampling.h:314-416
//   JitteredSampling::sample_points     tiling/Sampling.h:598-759
// Cells are runs of equal (key >> cell_shift) inside a node.  Per cell the reference selects
// std::min_element of squaredDistanceTo(target) = the FIRST minimum in Morton order.  Here every
// point evaluates its distance, a segmented inclusive min-scan (ties keep the earlier point) runs
// over the list in one pass (decoupled look-back carries the open cell across tiles), and the last
// point of each cell marks the winner in sel[].
#include "jitter_tables.cuh"

struct ArgminVal
{
  double d;
  u32 pos; // position in the input list
};

__device__ __forceinline__ ArgminVal
argmin_op(const ArgminVal& earlier, const ArgminVal& later)
{
  return (later.d < earlier.d) ? later : earlier; // strict: ties keep the earlier point
}

// get_octant_bounds recurrence, tiling/OctreeAlgorithms.cpp:3-18, applied `depth` times from the
// root (get_bounds_from_morton_index, OctreeAlgorithms.h:104-116).  ext/2 is exact.
__device__ __forceinline__ void
bounds_from_key(u64 key, int depth, const SwBounds& b, double mn[3], double mx[3])
{
  mn[0] = b.min[0];
  mn[1] = b.min[1];
  mn[2] = b.min[2];
  mx[0] = b.max[0];
  mx[1] = b.max[1];
  mx[2] = b.max[2];
  for (int level = 0; level < depth; ++level) {
    const u32 oct = (u32)(key >> (3 * (20 - level))) & 7u;
    const double hx = (mx[0] - mn[0]) * 0.5;
    const double hy = (mx[1] - mn[1]) * 0.5;
    const double hz = (mx[2] - mn[2]) * 0.5;
    if (oct & 4u)
      mn[0] = mn[0] + hx;
    if (oct & 2u)
      mn[1] = mn[1] + hy;
    if (oct & 1u)
      mn[2] = mn[2] + hz;
    mx[0] = mn[0] + hx;
    mx[1] = mn[1] + hy;
    mx[2] = mn[2] + hz;
  }
}

__device__ __forceinline__ double
squared_distance(const double p[3], const double t[3])
{
  // Vector3::squaredDistanceTo: (p - t).squaredLength() = x*x + y*y + z*z (math/Vector3.h:55-62);
  // compiled with -fmad=false so each product and sum is rounded separately.
  const double dx = p[0] - t[0];
  const double dy = p[1] - t[1];
  const double dz = p[2] - t[2];
  return dx * dx + dy * dy + dz * dz;
}

struct JitterNode
{
  int shift;  // key shift of the permutation grid cells
  int levels; // log2(cells per axis)
  u32 cells;
  double node_min[3];
  double grid_cell_size;
  double permutation_cell_size;
};

// per-node quantities of JitteredSampling::sample_points, Sampling.h:621-660
__device__ __forceinline__ u32
jitter_node_setup(u64 key, const SwArgminArgs& a, JitterNode& jn)
{
  double mn[3], mx[3];
  bounds_from_key(key, a.node_level + 1, a.bounds, mn, mx);
  const double ext_x = mx[0] - mn[0];
  const double perfect = ext_x / a.spacing_at_node;
  u32 x = __double2uint_rz(perfect);
  x |= x >> 1;
  x |= x >> 2;
  x |= x >> 4;
  x |= x >> 8;
  x |= x >> 16;
  const u32 cells = x - (x >> 1); // get_prev_power_of_two, util/stuff.cpp:340-349
  u32 err = 0;
  if (cells < 16)
    err = SW_ERR_JITTER_GRID_TOO_SMALL;
  const int levels = cells ? (31 - __clz(cells)) : 0;
  const int grid_level = a.node_level + levels;
  if (!err && grid_level >= 21)
    err = SW_ERR_JITTER_NODE_TOO_SMALL;
  jn.levels = levels;
  jn.cells = cells ? cells : 1;
  jn.shift = err ? 0 : 3 * (20 - grid_level);
  jn.node_min[0] = mn[0];
  jn.node_min[1] = mn[1];
  jn.node_min[2] = mn[2];
  jn.grid_cell_size = ext_x / (double)jn.cells;
  jn.permutation_cell_size = jn.grid_cell_size / (double)jn.cells;
  return err;
}

__device__ __forceinline__ void
jitter_target(u64 key, const SwArgminArgs& a, const JitterNode& jn, double t[3])
{
  const u64 rel = key >> jn.shift;
  const u64 grid_mask = (1ull << (3 * jn.levels)) - 1ull;
  const u64 idx = rel & grid_mask;
  const u64 lmask = (1ull << jn.levels) - 1ull;
  const u64 gz = contract_bits_by_3(idx) & lmask; // OctreeNodeIndex::to_grid_index, OctreeNodeIndex.h:357-363
  const u64 gy = contract_bits_by_3(idx >> 1) & lmask;
  const u64 gx = contract_bits_by_3(idx >> 2) & lmask;
  const u32 start_index = (3u * (u32)(a.node_level + 1)) % 16u;
  const u32 t0 = start_index, t1 = (start_index + 1) % 16u, t2 = (start_index + 2) % 16u;
  const u32 len = jn.cells < 64u ? jn.cells : 64u;
  const u32 ix = (u32)((gy + gz) % len), iy = (u32)((gx + gz) % len), iz = (u32)((gx + gy) % len);
  u32 px, py, pz;
  if (jn.cells <= 16u) {
    px = PERMUTATIONS_16[t0][ix];
    py = PERMUTATIONS_16[t1][iy];
    pz = PERMUTATIONS_16[t2][iz];
  } else if (jn.cells <= 32u) {
    px = PERMUTATIONS_32[t0][ix];
    py = PERMUTATIONS_32[t1][iy];
    pz = PERMUTATIONS_32[t2][iz];
  } else {
    px = PERMUTATIONS_64[t0][ix];
    py = PERMUTATIONS_64[t1][iy];
    pz = PERMUTATIONS_64[t2][iz];
  }
  px -= 1;
  py -= 1;
  pz -= 1;
  // target = node_min + (g * cell + p * sub), Sampling.h:735-739 (mul, mul, add, add: no FMA)
  t[0] = jn.node_min[0] + ((double)gx * jn.grid_cell_size + (double)px * jn.permutation_cell_size);
  t[1] = jn.node_min[1] + ((double)gy * jn.grid_cell_size + (double)py * jn.permutation_cell_size);
  t[2] = jn.node_min[2] + ((double)gz * jn.grid_cell_size + (double)pz * jn.permutation_cell_size);
}

// tile descriptors of the segmented scan: flag word + two 16-byte payload slots per tile
struct ArgminDesc
{
  double d;
  u32 pos;
  u32 has_head;
};

__global__ void __launch_bounds__(SWP_THREADS)
select_argmin_kernel(SwArgminArgs a, u64* __restrict__ status, u32* __restrict__ ticket)
{
  __shared__ u32 s_slot;
  __shared__ u32 s_w[SWP_WARPS];
  __shared__ ArgminDesc s_wagg[SWP_WARPS];
  __shared__ ArgminDesc s_tile_carry;
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 tile = take_ticket(ticket, &s_slot);
  const u64 base = (u64)tile * SW_SWEEP_TILE;
  const u32 lt = lanemask_lt();
  const size_t n_tiles = (size_t)((a.count + SW_SWEEP_TILE - 1) / SW_SWEEP_TILE);
  u32* flags = reinterpret_cast<u32*>(status);                            // n_tiles u32 (padded to u64)
  ArgminDesc* agg = reinterpret_cast<ArgminDesc*>(status + n_tiles);      // n_tiles x 16 B
  ArgminDesc* pfx = reinterpret_cast<ArgminDesc*>(status + 3 * n_tiles);  // n_tiles x 16 B

  // ---- phase 1: keys and node heads ------------------------------------------------------------
  u64 key[SWP_ITEMS];
  u32 nmask[SWP_ITEMS];
  u32 wheads = 0;
#pragma unroll
  for (int j = 0; j < SWP_ITEMS; ++j) {
    const u64 i = base + item_pos(warp, lane, j);
    bool nh = false;
    key[j] = 0;
    if (i < a.count) {
      key[j] = a.in_key[i] & SW_KEY_MASK;
      nh = (i == 0) || ((key[j] >> a.node_shift) != ((a.in_key[i - 1] & SW_KEY_MASK) >> a.node_shift));
    }
    nmask[j] = __ballot_sync(0xffffffffu, nh);
    wheads += __popc(nmask[j]);
  }
  u32 heads_total;
  const u32 hexcl = warp_totals_exclusive(wheads, warp, lane, s_w, heads_total);

  // ---- phase 2: per point head / tail flags and distance ------------------------------------------
  ArgminVal val[SWP_ITEMS];
  u32 hbits = 0, tbits = 0; // per item: this lane's element is a segment head / tail
  u32 local_err = 0;
  {
    u32 run = a.tile_rank0[tile] + hexcl;
#pragma unroll
    for (int j = 0; j < SWP_ITEMS; ++j) {
      const u64 i = base + item_pos(warp, lane, j);
      const u32 node_rank = run + __popc(nmask[j] & (lt | (1u << lane))) - 1;
      run += __popc(nmask[j]);
      val[j].d = 0.0;
      val[j].pos = (u32)i;
      bool head = true, tail = true;
      if (i < a.count) {
        bool active = true;
        if (a.allow_take_all) {
          const u32 cnt = a.node_start[node_rank + 1] - a.node_start[node_rank];
          active = (u64)cnt > a.max_points_per_node;
        }
        if (active) {
          const u64 k = key[j];
          const u32 idx = a.in_idx ? a.in_idx[i] : (u32)i;
          const double p[3] = { a.pos_sorted[3 * (u64)idx], a.pos_sorted[3 * (u64)idx + 1],
                                a.pos_sorted[3 * (u64)idx + 2] };
          double t[3];
          int cshift;
          if (a.sampling == SW_JITTERED) {
            JitterNode jn;
            const u32 e = jitter_node_setup(k, a, jn);
            if (e)
              local_err = local_err ? local_err : e;
            cshift = jn.shift;
            jitter_target(k, a, jn, t);
          } else {
            cshift = a.cell_shift;
            double mn[3], mx[3];
            bounds_from_key(k, a.cand_level + 1, a.bounds, mn, mx);
            // AABB::getCenter = min + extent()/2 (math/AABB.h:70)
            t[0] = mn[0] + (mx[0] - mn[0]) * 0.5;
            t[1] = mn[1] + (mx[1] - mn[1]) * 0.5;
            t[2] = mn[2] + (mx[2] - mn[2]) * 0.5;
          }
          val[j].d = squared_distance(p, t);
          const bool nh = (nmask[j] >> lane) & 1u;
          if (!nh) {
            const u64 pk = a.in_key[i - 1] & SW_KEY_MASK;
            head = (k >> cshift) != (pk >> cshift);
          }
          if (i + 1 < a.count) {
            const u64 nk = a.in_key[i + 1] & SW_KEY_MASK;
            tail = ((nk >> a.node_shift) != (k >> a.node_shift)) || ((nk >> cshift) != (k >> cshift));
          }
        }
      }
      hbits |= (head ? 1u : 0u) << j;
      tbits |= (tail ? 1u : 0u) << j;
    }
  }
  if (local_err)
    atomicMax(a.error_flag, local_err);

  // ---- phase 3: segmented inclusive min-scan inside the warp (items in order, lanes in order) -----
  // seen[j]: a head exists between the start of the warp's range and this element (inclusive)
  u32 seen_bits = 0;
  ArgminVal carry;
  carry.d = 0.0;
  carry.pos = 0;
  bool carry_valid = false; // false until the warp has processed its first element
  bool warp_seen = false;
#pragma unroll
  for (int j = 0; j < SWP_ITEMS; ++j) {
    ArgminVal v = val[j];
    bool f = (hbits >> j) & 1u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      ArgminVal up;
      up.d = __shfl_up_sync(0xffffffffu, v.d, o);
      up.pos = __shfl_up_sync(0xffffffffu, v.pos, o);
      const bool fup = __shfl_up_sync(0xffffffffu, (int)f, o) != 0;
      if (lane >= (u32)o) {
        if (!f)
          v = argmin_op(up, v);
        f = f || fup;
      }
    }
    // lanes before the first head of this item continue the run of the previous item
    if (!f && carry_valid)
      v = argmin_op(carry, v);
    const bool seen = f || warp_seen;
    seen_bits |= (seen ? 1u : 0u) << j;
    val[j] = v;
    // carry for the next item = value at lane 31
    carry.d = __shfl_sync(0xffffffffu, v.d, 31);
    carry.pos = __shfl_sync(0xffffffffu, v.pos, 31);
    carry_valid = true;
    warp_seen = __shfl_sync(0xffffffffu, (int)seen, 31) != 0;
  }
  if (lane == 31) {
    s_wagg[warp].d = carry.d;
    s_wagg[warp].pos = carry.pos;
    s_wagg[warp].has_head = warp_seen ? 1u : 0u;
  }
  __syncthreads();

  // ---- tile aggregate, look-back for the run that is open at the tile start ----------------------
  if (threadIdx.x == 0) {
    ArgminVal tv;
    tv.d = s_wagg[0].d;
    tv.pos = s_wagg[0].pos;
    bool th = s_wagg[0].has_head != 0;
#pragma unroll
    for (int w = 1; w < SWP_WARPS; ++w) {
      ArgminVal wv;
      wv.d = s_wagg[w].d;
      wv.pos = s_wagg[w].pos;
      if (s_wagg[w].has_head) {
        tv = wv;
        th = true;
      } else {
        tv = argmin_op(tv, wv);
      }
    }
    ArgminDesc mine;
    mine.d = tv.d;
    mine.pos = tv.pos;
    mine.has_head = th ? 1u : 0u;
    ArgminDesc carry_in;
    carry_in.d = 0.0;
    carry_in.pos = 0;
    carry_in.has_head = 0; // has_head == 0 here means "no carry" (tile 0 or cut by a head)
    bool have_carry = false;
    if (tile == 0) {
      pfx[0] = mine;
      __threadfence();
      st_relaxed_u32(flags, 2u);
    } else {
      agg[tile] = mine;
      __threadfence();
      st_relaxed_u32(flags + tile, 1u);
      ArgminVal acc;
      acc.d = 0.0;
      acc.pos = 0;
      long long t = (long long)tile - 1;
      while (t >= 0) {
        u32 f;
        do {
          f = ld_relaxed_u32(flags + t);
        } while (f == 0);
        __threadfence();
        const volatile ArgminDesc* src = (f == 2u) ? (pfx + t) : (agg + t);
        ArgminVal e;
        e.d = src->d;
        e.pos = src->pos;
        const u32 hh = src->has_head;
        acc = have_carry ? argmin_op(e, acc) : e;
        have_carry = true;
        if (f == 2u || hh)
          break;
        --t;
      }
      ArgminDesc incl = mine;
      if (!th && have_carry) {
        ArgminVal m;
        m.d = mine.d;
        m.pos = mine.pos;
        const ArgminVal r = argmin_op(acc, m);
        incl.d = r.d;
        incl.pos = r.pos;
      }
      // the inclusive prefix always describes a run that may continue: keep has_head as "valid"
      incl.has_head = 1u;
      pfx[tile] = incl;
      __threadfence();
      st_relaxed_u32(flags + tile, 2u);
      carry_in.d = acc.d;
      carry_in.pos = acc.pos;
    }
    carry_in.has_head = have_carry ? 1u : 0u;
    s_tile_carry = carry_in;
  }
  __syncthreads();

  // ---- carry into this warp = (tile carry, previous warps), then winners -------------------------
  ArgminVal cin;
  cin.d = s_tile_carry.d;
  cin.pos = s_tile_carry.pos;
  bool cin_valid = s_tile_carry.has_head != 0;
  for (u32 w = 0; w < warp; ++w) {
    ArgminVal wv;
    wv.d = s_wagg[w].d;
    wv.pos = s_wagg[w].pos;
    if (s_wagg[w].has_head || !cin_valid)
      cin = wv;
    else
      cin = argmin_op(cin, wv);
    cin_valid = true;
  }
#pragma unroll
  for (int j = 0; j < SWP_ITEMS; ++j) {
    const u64 i = base + item_pos(warp, lane, j);
    if (i < a.count && ((tbits >> j) & 1u)) {
      ArgminVal v = val[j];
      if (!((seen_bits >> j) & 1u) && cin_valid)
        v = argmin_op(cin, v);
      a.sel[v.pos] = 1;
    }
  }
}

void
launch_select_argmin(const SwArgminArgs& a, u64* status, u32* ticket, cudaStream_t stream)
{
  const size_t tiles = sweep_tiles(a.count);
  cudaMemsetAsync(status, 0, tiles * sizeof(u64), stream); // flag words
  cudaMemsetAsync(ticket, 0, sizeof(u32), stream);
  select_argmin_kernel<<<(u32)tiles, SWP_THREADS, 0, stream>>>(a, status, ticket);
}

// =============================================================================================
// K9  MIN_DISTANCE (placeholder until the conflict-resolution kernels land)
// =============================================================================================
cudaError_t
run_min_distance(const SwMinDistArgs&, const SwMinDistScratch&, cudaStream_t, u32*, u32*, u64*)
{
  return cudaErrorNotSupported;
}
