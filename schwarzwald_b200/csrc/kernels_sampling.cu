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

#define SWP_WARPS (SWP_THREADS / 32)

size_t
sweep_tiles(u64 count)
{
  const size_t t = (size_t)((count + SW_SWEEP_TILE - 1) / SW_SWEEP_TILE);
  return t ? t : 1;
}

// Points of node `rank` for the take-all decision (Sampling.h:201-208).  Single GPU: the length of
// the node's run.  Sharded: nodes above the shard prefix depth span GPUs, gcount holds the
// all-reduced count.
__device__ __forceinline__ u32
node_point_count(const u32* __restrict__ node_start, const u32* __restrict__ gcount, u32 rank)
{
  return gcount ? gcount[rank] : node_start[rank + 1] - node_start[rank];
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
// Two passes without any inter-CTA dependency.  A single-pass compaction needs a decoupled
// look-back chain; measured on B200 (tools/microbench/scan_chain.cu) the chain costs 3x the
// streaming time of the same kernel (0.49 ms vs 0.15 ms per 100 M keys, independent of tile size
// and window width), so reading the keys twice is the cheaper design:
//
//   level_count_kernel    [read 8 B/pt]  per element: node rank, take-all decision, selection
//                         flag; stores the flags as one bit per element, the number of selected
//                         points per tile and - for the points that stay - the point count of
//                         every child node (one atomic per run of equal child inside a warp)
//   level_scan_kernel     (one block) exclusive scan of the per-tile counts; exclusive scan of the
//                         child counts = node boundaries of the NEXT level, so no pass over the
//                         points is needed to find them (the reference does 8 linear scans per
//                         node, partition_points_into_child_octants, OctreeAlgorithms.h:240-265)
//   level_scatter_kernel  [read 12, write 12 B/pt]  moves selected points to the output chunk and
//                         the rest to the remainder list, both in order; node table rows
//
// Element position inside a tile is warp-striped (item_pos) in all three.
__device__ __forceinline__ void
warp_totals_exclusive2(u32 a_total, u32 b_total, u32 warp, u32 lane, u32* s_a, u32* s_b, u32& a_excl, u32& b_excl)
{
  if (lane == 0) {
    s_a[warp] = a_total;
    s_b[warp] = b_total;
  }
  __syncthreads();
  u32 ea = 0, eb = 0;
#pragma unroll
  for (int w = 0; w < SWP_WARPS; ++w) {
    ea += (w < (int)warp) ? s_a[w] : 0u;
    eb += (w < (int)warp) ? s_b[w] : 0u;
  }
  a_excl = ea;
  b_excl = eb;
}

// tile_rank0[t] = number of node heads before the first element of tile t
__global__ void __launch_bounds__(256)
tile_rank0_kernel(const u32* __restrict__ node_start, u32 n_nodes, u32 n_tiles, u32* __restrict__ tile_rank0)
{
  const u32 t = blockIdx.x * 256 + threadIdx.x;
  if (t >= n_tiles)
    return;
  const u64 base = (u64)t * SW_SWEEP_TILE;
  u32 lo = 0, hi = n_nodes; // first node whose start is >= base
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if ((u64)node_start[mid] < base)
      lo = mid + 1;
    else
      hi = mid;
  }
  tile_rank0[t] = lo;
}

void
launch_tile_rank0(const u32* node_start, u32 n_nodes, u64 count, u32* tile_rank0, cudaStream_t stream)
{
  const u32 tiles = (u32)sweep_tiles(count);
  tile_rank0_kernel<<<(tiles + 255) / 256, 256, 0, stream>>>(node_start, n_nodes, tiles, tile_rank0);
}

// the root node: one node holding every point
__global__ void
root_node_kernel(u32* node_start, u32 count)
{
  node_start[0] = 0;
  node_start[1] = count;
}

void
launch_root_node(u32* node_start, u64 count, cudaStream_t stream)
{
  root_node_kernel<<<1, 1, 0, stream>>>(node_start, (u32)count);
}

// ---- blocked tile layout of the count / scatter kernels -----------------------------------------
// Thread t owns the 8 consecutive elements [8t, 8t + 8) of its tile, so node / cell heads, selection
// bits and child runs are thread-local bit operations and the warp-wide collectives (a ballot costs
// ~2.6 SM-cycles, a popc ~2.3 on B200) are needed once per 256 elements instead of once per 32.
// Global loads and stores stay coalesced: tiles are transposed through shared memory, padded by one
// slot per 8 so that the 72-byte thread stride is conflict-free.
#define BLK_ITEMS 8
#define BLK_PAD(p) ((p) + ((p) >> 3))
#define BLK_SLOTS (SW_SWEEP_TILE + SW_SWEEP_TILE / 8)
static_assert(SWP_ITEMS == BLK_ITEMS, "blocked layout assumes 8 elements per thread");

// keys of the tile into registers (blocked) + the key before the thread's first element
__device__ __forceinline__ void
load_keys_blocked(const u64* __restrict__ in_key, u64 base, u64 count, u64* s_k, u64* s_prev, u64 k[BLK_ITEMS], u64& prev)
{
  const u32 tid = threadIdx.x;
  if (base + SW_SWEEP_TILE <= count) { // all tiles but the last: no bounds tests
    const u64* __restrict__ kp = in_key + base;
#pragma unroll
    for (int j = 0; j < BLK_ITEMS; ++j) {
      const u32 p = j * SWP_THREADS + tid;
      s_k[BLK_PAD(p)] = kp[p] & SW_KEY_MASK;
    }
  } else {
#pragma unroll
    for (int j = 0; j < BLK_ITEMS; ++j) {
      const u32 p = j * SWP_THREADS + tid;
      const u64 i = base + p;
      s_k[BLK_PAD(p)] = (i < count) ? (in_key[i] & SW_KEY_MASK) : ~0ull;
    }
  }
  if (tid == 0)
    *s_prev = base ? (in_key[base - 1] & SW_KEY_MASK) : 0ull;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j)
    k[j] = s_k[9 * tid + j];
  prev = tid ? s_k[9 * tid - 2] : *s_prev; // BLK_PAD(8 * tid - 1)
}

// Head bits of the thread's elements: bit j of nh (ch) is set when element j starts a new run of
// key >> node_shift (key >> cell_shift).  Prefixes differ iff the XOR of the keys has a bit at or
// above the shift, so both tests share one XOR per element.
__device__ __forceinline__ void
head_bits2(const u64 k[BLK_ITEMS], u64 prev, int node_shift, int cell_shift, bool first_of_list, u32& nh, u32& ch)
{
  const u64 nmask = ~0ull << node_shift, cmask = ~0ull << cell_shift; // shifts are <= 63
  u32 n = 0, c = 0;
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j) {
    const u64 x = k[j] ^ (j ? k[j - 1] : prev);
    n |= ((x & nmask) ? 1u : 0u) << j;
    c |= ((x & cmask) ? 1u : 0u) << j;
  }
  if (first_of_list) {
    n |= 1u;
    c |= 1u;
  }
  nh = n;
  ch = c | n;
}

// exclusive block scan of a packed pair of counters (each total < 2^16); returns the block total
__device__ __forceinline__ u32
block_scan_packed(u32 v, u32* s_w, u32& excl)
{
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32)o)
      incl += up;
  }
  if (lane == 31)
    s_w[warp] = incl;
  __syncthreads();
  u32 wofs = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SWP_WARPS; ++w) {
    const u32 x = s_w[w];
    wofs += (w < (int)warp) ? x : 0u;
    tot += x;
  }
  excl = wofs + incl - v;
  return tot;
}

#define CHILD_WINDOW 256
#ifndef COUNT_MIN_CTAS
#define COUNT_MIN_CTAS 6 /* 40 registers, no spills; 5 / 8: within 1 % */
#endif
__global__ void __launch_bounds__(SWP_THREADS, COUNT_MIN_CTAS)
level_count_kernel(SwLevelArgs a)
{
  __shared__ u64 s_k[BLK_SLOTS];
  __shared__ u64 s_prev;
  __shared__ u32 s_w[SWP_WARPS];
  __shared__ u32 s_w2[SWP_WARPS];
  __shared__ u32 s_child[CHILD_WINDOW]; // child counters of the first CHILD_WINDOW slots touched by the tile
  const u32 tid = threadIdx.x, lane = tid & 31;
  const u32 tile = blockIdx.x;
  const u64 base = (u64)tile * SW_SWEEP_TILE;
  const u64 e0 = base + 8ull * tid;
  for (u32 i = tid; i < CHILD_WINDOW; i += SWP_THREADS)
    s_child[i] = 0;

  u64 k[BLK_ITEMS];
  u64 prev;
  load_keys_blocked(a.in_key, base, a.count, s_k, &s_prev, k, prev);
  const u32 nvalid = e0 >= a.count ? 0u : (a.count - e0 < 8 ? (u32)(a.count - e0) : 8u);
  const u32 valid = (1u << nvalid) - 1u;

  // ---- node heads, cell heads --------------------------------------------------------------------
  u32 nh, ch;
  head_bits2(k, prev, a.node_shift, a.cell_shift, e0 == 0, nh, ch);
  nh &= valid;
  ch &= valid;
  u32 hexcl;
  block_scan_packed(__popc(nh), s_w, hexcl);

  // ---- node rank -> take-all decision -> selection bits ----------------------------------------------
  const u32 rank0 = a.tile_rank0[tile];
  u32 rank = rank0 + hexcl - 1u; // node of the element before my first one (0xFFFFFFFF: none yet)
  bool take = a.force_all != 0;
  if (!take && a.allow_take_all && nvalid && !(nh & 1u))
    take = (u64)node_point_count(a.node_start, a.node_gcount, rank) <= a.max_points_per_node;
  u32 other = 0; // strategy flags of the 8 elements (one byte each in a.sel)
  if (a.sampling != SW_RANDOM_GRID && !a.force_all && nvalid) {
    if (nvalid == 8) {
      const u64 bytes = *reinterpret_cast<const u64*>(a.sel + e0);
#pragma unroll
      for (int j = 0; j < BLK_ITEMS; ++j)
        other |= (((bytes >> (8 * j)) & 0xFFu) == 1u ? 1u : 0u) << j;
    } else {
      for (u32 j = 0; j < nvalid; ++j)
        other |= (a.sel[e0 + j] == 1 ? 1u : 0u) << j;
    }
  }
  const u32 pick = (a.sampling == SW_RANDOM_GRID) ? ch : other; // Sampling.h:253-284: first point of each cell run
  u32 sel = 0;
  u32 slot0 = 0;          // child slot of my first element
  bool one_slot = true;   // all my elements fall into the same child of the same node
  const int child_shift = a.node_shift - 3;
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j) {
    if ((nh >> j) & 1u) {
      ++rank;
      if (!a.force_all && a.allow_take_all)
        take = (u64)node_point_count(a.node_start, a.node_gcount, rank) <= a.max_points_per_node;
    }
    sel |= ((take ? 1u : ((pick >> j) & 1u)) << j);
    if (a.child_count) {
      const u32 slot = rank * 8u + (u32)((k[j] >> child_shift) & 7u);
      if (j == 0)
        slot0 = slot;
      else if (((valid >> j) & 1u) && slot != slot0)
        one_slot = false;
    }
  }
  sel &= valid;
  reinterpret_cast<unsigned char*>(a.selbits)[(size_t)tile * SWP_THREADS + tid] = (unsigned char)sel;

  // ---- selected points of the tile ---------------------------------------------------------------------
  u32 sexcl;
  const u32 tile_selected = block_scan_packed(__popc(sel), s_w2, sexcl);
  if (tid == 0)
    a.tile_sel[tile] = tile_selected;

  // ---- points that stay, counted per child node ----------------------------------------------------------
  if (a.child_count) {
    const u32 slot_base = (rank0 ? rank0 - 1u : 0u) * 8u; // slots of this tile start here or later
    const u32 rem = valid & ~sel;
    const u32 lead_slot = __shfl_sync(0xffffffffu, slot0, 0);
    const bool uniform = __all_sync(0xffffffffu, nvalid == 0 || (one_slot && slot0 == lead_slot));
    if (uniform) { // the whole warp (256 points) lies in one child: one atomic
      u32 c = __popc(rem);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0 && c) {
        if (lead_slot - slot_base < CHILD_WINDOW)
          atomicAdd(&s_child[lead_slot - slot_base], c);
        else
          atomicAdd(&a.child_count[lead_slot], c);
      }
    } else { // runs of equal child inside the thread
      u32 r2 = rank0 + hexcl - 1u;
      u32 cur = 0xFFFFFFFFu, cnt = 0;
#pragma unroll
      for (int j = 0; j < BLK_ITEMS; ++j) {
        if ((nh >> j) & 1u)
          ++r2;
        if ((rem >> j) & 1u) {
          const u32 slot = r2 * 8u + (u32)((k[j] >> child_shift) & 7u);
          if (slot != cur) {
            if (cnt) {
              if (cur - slot_base < CHILD_WINDOW)
                atomicAdd(&s_child[cur - slot_base], cnt);
              else
                atomicAdd(&a.child_count[cur], cnt);
            }
            cur = slot;
            cnt = 0;
          }
          ++cnt;
        }
      }
      if (cnt) {
        if (cur - slot_base < CHILD_WINDOW)
          atomicAdd(&s_child[cur - slot_base], cnt);
        else
          atomicAdd(&a.child_count[cur], cnt);
      }
    }
    __syncthreads();
    for (u32 i = tid; i < CHILD_WINDOW; i += SWP_THREADS) {
      const u32 c = s_child[i];
      if (c)
        atomicAdd(&a.child_count[slot_base + i], c);
    }
  }
}

// One block.  (A) exclusive scan of tile_sel in place, total -> *n_selected.  (B) child counts ->
// node_start of the next level (start position in the remainder list of every non-empty child,
// plus the sentinel) and *n_nodes_next.
#define SCAN_THREADS 1024
__global__ void __launch_bounds__(SCAN_THREADS)
level_scan_kernel(u32* __restrict__ tile_sel, u32 n_tiles, u64* __restrict__ n_selected,
                  const u32* __restrict__ child_count, u32 n_slots, u32* __restrict__ next_node_start,
                  u32* __restrict__ n_nodes_next)
{
  // tile_sel == nullptr: only part (B) (the single-pass compaction has no tile counts to scan)
  __shared__ u32 s_wa[32], s_wb[32];
  __shared__ u32 s_carry_a, s_carry_b;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    s_carry_a = 0;
    s_carry_b = 0;
  }
  __syncthreads();
  // ---- (A) ---- (8 consecutive tiles per thread per iteration)
  for (u32 t0 = 0; tile_sel && t0 < n_tiles; t0 += SCAN_THREADS * 8) {
    const u32 t = t0 + threadIdx.x * 8;
    u32 v[8];
    u32 sum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[q] = (t + q < n_tiles) ? tile_sel[t + q] : 0u;
      sum += v[q];
    }
    u32 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (u32)o)
        incl += up;
    }
    if (lane == 31)
      s_wa[warp] = incl;
    __syncthreads();
    u32 wofs = 0;
    for (u32 w = 0; w < warp; ++w)
      wofs += s_wa[w];
    const u32 carry = s_carry_a;
    u32 run = carry + wofs + incl - sum;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (t + q < n_tiles)
        tile_sel[t + q] = run;
      run += v[q];
    }
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1)
      s_carry_a = carry + wofs + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && tile_sel)
    *n_selected = s_carry_a;
  // ---- (B) ----
  if (!child_count)
    return;
  __syncthreads();
  if (threadIdx.x == 0)
    s_carry_a = 0;
  __syncthreads();
  // 8 consecutive child slots per thread per iteration (the slots of one node): two scans in one, points and
  // non-empty children
  for (u32 t0 = 0; t0 < n_slots; t0 += SCAN_THREADS * 8) {
    const u32 t = t0 + threadIdx.x * 8;
    u32 v[8];
    u32 sa = 0, sb = 0;
    if (t + 8 <= n_slots) { // n_slots is a multiple of 8 and child_count is 16-byte aligned
      const uint4 lo4 = *reinterpret_cast<const uint4*>(child_count + t);
      const uint4 hi4 = *reinterpret_cast<const uint4*>(child_count + t + 4);
      v[0] = lo4.x; v[1] = lo4.y; v[2] = lo4.z; v[3] = lo4.w;
      v[4] = hi4.x; v[5] = hi4.y; v[6] = hi4.z; v[7] = hi4.w;
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        v[q] = (t + q < n_slots) ? child_count[t + q] : 0u;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sa += v[q];
      sb += v[q] ? 1u : 0u;
    }
    u32 ia = sa, ib = sb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 ua = __shfl_up_sync(0xffffffffu, ia, o);
      const u32 ub = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= (u32)o) {
        ia += ua;
        ib += ub;
      }
    }
    if (lane == 31) {
      s_wa[warp] = ia;
      s_wb[warp] = ib;
    }
    __syncthreads();
    u32 oa = 0, ob = 0;
    for (u32 w = 0; w < warp; ++w) {
      oa += s_wa[w];
      ob += s_wb[w];
    }
    const u32 ca = s_carry_a, cb = s_carry_b;
    u32 run_a = ca + oa + ia - sa, run_b = cb + ob + ib - sb;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (v[q]) {
        next_node_start[run_b] = run_a;
        ++run_b;
      }
      run_a += v[q];
    }
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1) {
      s_carry_a = ca + oa + ia;
      s_carry_b = cb + ob + ib;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    next_node_start[s_carry_b] = s_carry_a; // sentinel = points that stay
    *n_nodes_next = s_carry_b;
  }
}

#ifndef SCATTER_MIN_CTAS
#define SCATTER_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(SWP_THREADS, SCATTER_MIN_CTAS)
level_scatter_kernel(SwLevelArgs a)
{
  __shared__ u64 s_k[BLK_SLOTS];
  __shared__ u32 s_i[BLK_SLOTS];
  __shared__ u64 s_prev;
  __shared__ u32 s_w[SWP_WARPS];
  const u32 tid = threadIdx.x;
  const u32 tile = blockIdx.x;
  const u64 base = (u64)tile * SW_SWEEP_TILE;
  const u64 e0 = base + 8ull * tid;
  const u32 tile_valid = (a.count - base) < SW_SWEEP_TILE ? (u32)(a.count - base) : SW_SWEEP_TILE;

  // ---- load: keys and ids, coalesced, transposed to the blocked layout ----------------------------------
  if (a.in_idx) {
    const u32* __restrict__ ip = a.in_idx + base;
    if (tile_valid == SW_SWEEP_TILE) {
#pragma unroll
      for (int j = 0; j < BLK_ITEMS; ++j) {
        const u32 p = j * SWP_THREADS + tid;
        s_i[BLK_PAD(p)] = ip[p];
      }
    } else {
#pragma unroll
      for (int j = 0; j < BLK_ITEMS; ++j) {
        const u32 p = j * SWP_THREADS + tid;
        s_i[BLK_PAD(p)] = (p < tile_valid) ? ip[p] : 0u;
      }
    }
  }
  u64 k[BLK_ITEMS];
  u64 prev;
  load_keys_blocked(a.in_key, base, a.count, s_k, &s_prev, k, prev);
  u32 idx[BLK_ITEMS];
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j)
    idx[j] = a.in_idx ? s_i[9 * tid + j] : (u32)(e0 + j);
  const u32 nvalid = e0 >= a.count ? 0u : (a.count - e0 < 8 ? (u32)(a.count - e0) : 8u);
  const u32 valid = (1u << nvalid) - 1u;
  u32 nh, ch_unused;
  head_bits2(k, prev, a.node_shift, a.node_shift, e0 == 0, nh, ch_unused);
  nh &= valid;
  const u32 sel = reinterpret_cast<const unsigned char*>(a.selbits)[(size_t)tile * SWP_THREADS + tid];

  // ---- ranks inside the tile: node heads and selected points before my first element ----------------------
  u32 excl;
  const u32 tot = block_scan_packed((__popc(nh) << 16) | __popc(sel), s_w, excl); // also: s_k / s_i are free now
  const u32 hexcl = excl >> 16, sexcl = excl & 0xFFFFu;
  const u32 tile_selected = tot & 0xFFFFu;
  const u64 tile_off = a.tile_sel[tile]; // selected before this tile (exclusive scan)

  // ---- node table rows (from registers) and staging in output order -----------------------------------------
  u32 rank = a.tile_rank0[tile] + hexcl - 1u;
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j) {
    if ((valid >> j) & 1u) {
      const u32 sbefore = sexcl + __popc(sel & ((1u << j) - 1u));
      if ((nh >> j) & 1u) {
        ++rank;
        bool take = false;
        if (!a.force_all && a.allow_take_all)
          take = (u64)node_point_count(a.node_start, a.node_gcount, rank) <= a.max_points_per_node;
        a.node_index[a.node_base + rank] = k[j] >> a.node_shift;
        // bit 63 = SW_NODE_TAKE_ALL (decoded by the host when it builds the node table)
        a.node_first[a.node_base + rank] = (a.out_offset + tile_off + sbefore) | ((u64)(take ? 1u : 0u) << 63);
      }
      // selected points first, then the points that stay, both in input order
      const u32 p = ((sel >> j) & 1u) ? sbefore : tile_selected + (8u * tid + j - sbefore);
      s_k[BLK_PAD(p)] = k[j];
      s_i[BLK_PAD(p)] = idx[j];
    }
  }
  __syncthreads();

  // ---- coalesced copy-out: one contiguous range of the output chunk, one of the remainder list -------------
  u64* const out_key = a.out_key + a.out_offset + tile_off;
  u32* const out_idx = a.out_idx + a.out_offset + tile_off;
  const u64 rem_base = base - tile_off;
#pragma unroll
  for (int m = 0; m < BLK_ITEMS; ++m) {
    const u32 q = m * SWP_THREADS + tid;
    if (q < tile_selected) {
      out_key[q] = s_k[BLK_PAD(q)];
      out_idx[q] = s_i[BLK_PAD(q)];
    } else if (q < tile_valid && a.rem_key) {
      const u32 r = q - tile_selected;
      a.rem_key[rem_base + r] = s_k[BLK_PAD(q)];
      a.rem_idx[rem_base + r] = s_i[BLK_PAD(q)];
    }
  }
}

// ---- single-pass variant (opt-in: SWGPU_COMPACT=1pass) -----------------------------------------------------
// level_count + level_scatter in one kernel: the tile decides its selection, publishes the number of selected
// points, and while the decoupled look-back over the tile descriptors resolves (warp 0) the other warps count the
// children and stage the tile in output order.  One read of (key, id) and one write per point instead of a second
// read of the keys, no selection bits, no tile scan.  Result identical to the two-pass kernels (the GPU parity
// suite passes with it), but measured SLOWER on B200 in round 2: 0.88 ms instead of 0.82 ms per 100 M-point level
// (C2 sweep 3.86 vs 3.61 ms, C3 shape at 100 M 13.70 vs 13.23 ms), and slower still with LB_WIDE = 8 descriptors
// per lane and round trip (4.35 ms): the wait is not the walk over the descriptors but the block barrier behind the
// look-back warp, a few memory round trips per tile that the 4 resident blocks per SM do not cover.  Kept for A/B
// runs; the two-pass kernels stay the default.
#ifndef LB_WIDE
#define LB_WIDE 1
#endif
__device__ __forceinline__ u64
lookback_resolve(u64* status, u32 tile, u64 aggregate) // all lanes of one warp; the aggregate is already published
{
  const u32 lane = threadIdx.x & 31;
  if (tile == 0)
    return 0;
  u64 prefix = 0;
  long long pos = (long long)tile - 1;
  bool done = false;
  while (!done) {
    u64 s[LB_WIDE];
#pragma unroll
    for (int k = 0; k < LB_WIDE; ++k) {
      const long long idx = pos - k * 32 - lane;
      s[k] = (idx >= 0) ? ld_relaxed_u64(status + idx) : SW_LB_PFX; // tiles before 0: inclusive prefix 0
    }
#pragma unroll
    for (int k = 0; k < LB_WIDE; ++k) {
      if (!done) {
        const long long idx = pos - k * 32 - lane;
        u64 v = s[k];
        while ((v >> 62) == 0)
          v = ld_relaxed_u64(status + idx);
        const u32 pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first_p = pm ? (__ffs(pm) - 1) : 32;
        u64 c = ((int)lane <= first_p) ? (v & SW_LB_MASK) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          c += __shfl_xor_sync(0xffffffffu, c, o);
        prefix += c;
        done = pm != 0;
      }
    }
    pos -= 32 * LB_WIDE;
  }
  if (lane == 0)
    st_relaxed_u64(status + tile, SW_LB_PFX | ((prefix + aggregate) & SW_LB_MASK));
  return prefix;
}

__global__ void __launch_bounds__(SWP_THREADS, 4)
level_compact_fused_kernel(SwLevelArgs a, u64* __restrict__ n_selected)
{
  __shared__ u64 s_k[BLK_SLOTS];
  __shared__ u32 s_i[BLK_SLOTS];
  __shared__ u64 s_prev;
  __shared__ u64 s_prefix;
  __shared__ u32 s_w[SWP_WARPS];
  __shared__ u32 s_w2[SWP_WARPS];
  __shared__ u32 s_slot;
  __shared__ u32 s_child[CHILD_WINDOW];
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 tile = take_ticket(a.ticket, &s_slot);
  const u64 base = (u64)tile * SW_SWEEP_TILE;
  const u64 e0 = base + 8ull * tid;
  const u32 tile_valid = (a.count - base) < SW_SWEEP_TILE ? (u32)(a.count - base) : SW_SWEEP_TILE;
  for (u32 i = tid; i < CHILD_WINDOW; i += SWP_THREADS)
    s_child[i] = 0;

  // ---- load: keys and ids, coalesced, transposed to the blocked layout ----------------------------------
  if (a.in_idx) {
#pragma unroll
    for (int j = 0; j < BLK_ITEMS; ++j) {
      const u32 p = j * SWP_THREADS + tid;
      s_i[BLK_PAD(p)] = (p < tile_valid) ? a.in_idx[base + p] : 0u;
    }
  }
  u64 k[BLK_ITEMS];
  u64 prev;
  load_keys_blocked(a.in_key, base, a.count, s_k, &s_prev, k, prev);
  u32 idx[BLK_ITEMS];
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j)
    idx[j] = a.in_idx ? s_i[9 * tid + j] : (u32)(e0 + j);
  const u32 nvalid = e0 >= a.count ? 0u : (a.count - e0 < 8 ? (u32)(a.count - e0) : 8u);
  const u32 valid = (1u << nvalid) - 1u;

  // ---- node heads, cell heads, node rank -> take-all decision -> selection bits (as level_count_kernel) -----
  u32 nh, ch;
  head_bits2(k, prev, a.node_shift, a.cell_shift, e0 == 0, nh, ch);
  nh &= valid;
  ch &= valid;
  u32 hexcl;
  block_scan_packed(__popc(nh), s_w, hexcl); // (barrier: every thread holds its keys and ids in registers now)
  const u32 rank0 = a.tile_rank0[tile];
  const u32 rank_before = rank0 + hexcl - 1u; // node of the element before my first one (0xFFFFFFFF: none yet)
  u32 rank = rank_before;
  bool take = a.force_all != 0;
  if (!take && a.allow_take_all && nvalid && !(nh & 1u))
    take = (u64)node_point_count(a.node_start, a.node_gcount, rank) <= a.max_points_per_node;
  u32 other = 0; // strategy flags of the 8 elements (one byte each in a.sel)
  if (a.sampling != SW_RANDOM_GRID && !a.force_all && nvalid) {
    if (nvalid == 8) {
      const u64 bytes = *reinterpret_cast<const u64*>(a.sel + e0);
#pragma unroll
      for (int j = 0; j < BLK_ITEMS; ++j)
        other |= (((bytes >> (8 * j)) & 0xFFu) == 1u ? 1u : 0u) << j;
    } else {
      for (u32 j = 0; j < nvalid; ++j)
        other |= (a.sel[e0 + j] == 1 ? 1u : 0u) << j;
    }
  }
  const u32 pick = (a.sampling == SW_RANDOM_GRID) ? ch : other; // Sampling.h:253-284: first point of each cell run
  u32 sel = 0, takebits = 0;
  u32 slot0 = 0;        // child slot of my first element
  bool one_slot = true; // all my elements fall into the same child of the same node
  const int child_shift = a.node_shift - 3;
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j) {
    if ((nh >> j) & 1u) {
      ++rank;
      if (!a.force_all && a.allow_take_all)
        take = (u64)node_point_count(a.node_start, a.node_gcount, rank) <= a.max_points_per_node;
    }
    sel |= ((take ? 1u : ((pick >> j) & 1u)) << j);
    takebits |= (take ? 1u : 0u) << j;
    if (a.child_count) {
      const u32 slot = rank * 8u + (u32)((k[j] >> child_shift) & 7u);
      if (j == 0)
        slot0 = slot;
      else if (((valid >> j) & 1u) && slot != slot0)
        one_slot = false;
    }
  }
  sel &= valid;

  // ---- selected points of the tile: publish, then look back while the rest of the block goes on -------------
  u32 sexcl;
  const u32 tile_selected = block_scan_packed(__popc(sel), s_w2, sexcl);
  if (tid == 0)
    st_relaxed_u64(a.status + tile, (tile == 0 ? SW_LB_PFX : SW_LB_AGG) | (u64)tile_selected);

  // ---- points that stay, counted per child node (as level_count_kernel) -------------------------------------
  if (a.child_count) {
    const u32 slot_base = (rank0 ? rank0 - 1u : 0u) * 8u; // slots of this tile start here or later
    const u32 rem = valid & ~sel;
    const u32 lead_slot = __shfl_sync(0xffffffffu, slot0, 0);
    const bool uniform = __all_sync(0xffffffffu, nvalid == 0 || (one_slot && slot0 == lead_slot));
    if (uniform) { // the whole warp (256 points) lies in one child: one atomic
      u32 c = __popc(rem);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0 && c) {
        if (lead_slot - slot_base < CHILD_WINDOW)
          atomicAdd(&s_child[lead_slot - slot_base], c);
        else
          atomicAdd(&a.child_count[lead_slot], c);
      }
    } else { // runs of equal child inside the thread
      u32 r2 = rank_before;
      u32 cur = 0xFFFFFFFFu, cnt = 0;
#pragma unroll
      for (int j = 0; j < BLK_ITEMS; ++j) {
        if ((nh >> j) & 1u)
          ++r2;
        if ((rem >> j) & 1u) {
          const u32 slot = r2 * 8u + (u32)((k[j] >> child_shift) & 7u);
          if (slot != cur) {
            if (cnt) {
              if (cur - slot_base < CHILD_WINDOW)
                atomicAdd(&s_child[cur - slot_base], cnt);
              else
                atomicAdd(&a.child_count[cur], cnt);
            }
            cur = slot;
            cnt = 0;
          }
          ++cnt;
        }
      }
      if (cnt) {
        if (cur - slot_base < CHILD_WINDOW)
          atomicAdd(&s_child[cur - slot_base], cnt);
        else
          atomicAdd(&a.child_count[cur], cnt);
      }
    }
  }

  // ---- staging in output order: selected points first, then the points that stay, both in input order --------
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j) {
    if ((valid >> j) & 1u) {
      const u32 sbefore = sexcl + __popc(sel & ((1u << j) - 1u));
      const u32 p = ((sel >> j) & 1u) ? sbefore : tile_selected + (8u * tid + j - sbefore);
      s_k[BLK_PAD(p)] = k[j];
      s_i[BLK_PAD(p)] = idx[j];
    }
  }
  if (warp == 0) {
    const u64 pfx = lookback_resolve(a.status, tile, (u64)tile_selected);
    if (lane == 0)
      s_prefix = pfx;
  }
  __syncthreads();
  const u64 tile_off = s_prefix; // selected before this tile
  if (a.child_count) {
    const u32 slot_base = (rank0 ? rank0 - 1u : 0u) * 8u;
    for (u32 i = tid; i < CHILD_WINDOW; i += SWP_THREADS) {
      const u32 c = s_child[i];
      if (c)
        atomicAdd(&a.child_count[slot_base + i], c);
    }
  }
  if (tid == 0 && base + SW_SWEEP_TILE >= a.count)
    *n_selected = tile_off + tile_selected;

  // ---- node table rows ---------------------------------------------------------------------------------------
  if (nh) {
    u32 r3 = rank_before;
#pragma unroll
    for (int j = 0; j < BLK_ITEMS; ++j) {
      if ((nh >> j) & 1u) {
        ++r3;
        const u32 sbefore = sexcl + __popc(sel & ((1u << j) - 1u));
        const bool tk = !a.force_all && a.allow_take_all && ((takebits >> j) & 1u);
        a.node_index[a.node_base + r3] = k[j] >> a.node_shift;
        // bit 63 = SW_NODE_TAKE_ALL (decoded by the host when it builds the node table)
        a.node_first[a.node_base + r3] = (a.out_offset + tile_off + sbefore) | ((u64)(tk ? 1u : 0u) << 63);
      }
    }
  }

  // ---- coalesced copy-out: one contiguous range of the output chunk, one of the remainder list -------------
  u64* const out_key = a.out_key + a.out_offset + tile_off;
  u32* const out_idx = a.out_idx + a.out_offset + tile_off;
  const u64 rem_base = base - tile_off;
#pragma unroll
  for (int m = 0; m < BLK_ITEMS; ++m) {
    const u32 q = m * SWP_THREADS + tid;
    if (q < tile_selected) {
      out_key[q] = s_k[BLK_PAD(q)];
      out_idx[q] = s_i[BLK_PAD(q)];
    } else if (q < tile_valid && a.rem_key) {
      const u32 r = q - tile_selected;
      a.rem_key[rem_base + r] = s_k[BLK_PAD(q)];
      a.rem_idx[rem_base + r] = s_i[BLK_PAD(q)];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// node tables that come from per-node data instead of a pass over the points
// ---------------------------------------------------------------------------------------------
// One block: stable compaction of the entries with flag != 0: out[rank] = value, out[total] = sentinel.
template<typename Flag, typename Value>
__device__ __forceinline__ void
block_compact(u32 n, Flag flag_of, Value value_of, u32* __restrict__ out, u32 sentinel, u32* __restrict__ n_out)
{
  __shared__ u32 s_wf[32];
  __shared__ u32 s_carry;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    s_carry = 0;
  __syncthreads();
  for (u32 t0 = 0; t0 < n; t0 += SCAN_THREADS) {
    const u32 t = t0 + threadIdx.x;
    const u32 f = (t < n && flag_of(t)) ? 1u : 0u;
    const u32 m = __ballot_sync(0xffffffffu, f);
    if (lane == 0)
      s_wf[warp] = __popc(m);
    __syncthreads();
    u32 wofs = 0, tot = 0;
    for (u32 w = 0; w < SCAN_THREADS / 32; ++w) {
      wofs += (w < warp) ? s_wf[w] : 0u;
      tot += s_wf[w];
    }
    const u32 carry = s_carry;
    if (f)
      out[carry + wofs + __popc(m & lanemask_lt())] = value_of(t);
    __syncthreads();
    if (threadIdx.x == 0)
      s_carry = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[s_carry] = sentinel;
    *n_out = s_carry;
  }
}

// FAST start nodes (split_indexed_points_into_subranges, TilingAlgorithms.cpp:1537-1578): the
// non-empty prefixes of S levels, from the level-5 bin boundaries of the sorted keys.
__global__ void __launch_bounds__(SCAN_THREADS)
start_nodes_kernel(const u32* __restrict__ bin_start, int start_levels, u32 count, u32* __restrict__ node_start,
                   u32* __restrict__ n_nodes)
{
  const u32 group = 1u << (3 * (6 - start_levels));
  const u32 n_groups = 262144u / group;
  block_compact(
    n_groups, [&](u32 g) { return bin_start[(g + 1) * group] > bin_start[g * group]; },
    [&](u32 g) { return bin_start[g * group]; }, node_start, count, n_nodes);
}

// Reconstruct (reconstruct_single_node, TilingAlgorithms.cpp:1661-1715): the input of a parent is
// the concatenation of its children's stored points = a run of child nodes in the previous chunk.
__global__ void __launch_bounds__(SCAN_THREADS)
parent_nodes_kernel(const u64* __restrict__ child_index, const u64* __restrict__ child_first, u32 n_children,
                    u64 chunk_offset, u32 chunk_count, u32* __restrict__ node_start, u32* __restrict__ n_nodes)
{
  block_compact(
    n_children, [&](u32 c) { return c == 0 || (child_index[c] >> 3) != (child_index[c - 1] >> 3); },
    [&](u32 c) { return (u32)((child_first[c] & ~(1ull << 63)) - chunk_offset); }, node_start, chunk_count, n_nodes);
}

void
launch_start_nodes(const u32* bin_start, int start_levels, u64 count, u32* node_start, u32* n_nodes, cudaStream_t stream)
{
  start_nodes_kernel<<<1, SCAN_THREADS, 0, stream>>>(bin_start, start_levels, (u32)count, node_start, n_nodes);
}

void
launch_parent_nodes(const u64* child_index, const u64* child_first, u32 n_children, u64 chunk_offset, u64 chunk_count,
                    u32* node_start, u32* n_nodes, cudaStream_t stream)
{
  parent_nodes_kernel<<<1, SCAN_THREADS, 0, stream>>>(child_index, child_first, n_children, chunk_offset,
                                                       (u32)chunk_count, node_start, n_nodes);
}

void
launch_level_compact(const SwLevelArgs& a, u64* n_selected, u32 n_child_slots, u32* next_node_start, u32* n_nodes_next,
                     cudaStream_t stream)
{
  const u32 tiles = (u32)sweep_tiles(a.count);
  if (a.child_count)
    cudaMemsetAsync(a.child_count, 0, (size_t)n_child_slots * sizeof(u32), stream);
  if (a.status) { // single pass (a.status == nullptr selects the two-pass kernels: A/B measurements)
    cudaMemsetAsync(a.status, 0, (size_t)tiles * sizeof(u64), stream);
    cudaMemsetAsync(a.ticket, 0, sizeof(u32), stream);
    level_compact_fused_kernel<<<tiles, SWP_THREADS, 0, stream>>>(a, n_selected);
    if (a.child_count)
      level_scan_kernel<<<1, SCAN_THREADS, 0, stream>>>(nullptr, 0, n_selected, a.child_count, n_child_slots,
                                                        next_node_start, n_nodes_next);
    return;
  }
  level_count_kernel<<<tiles, SWP_THREADS, 0, stream>>>(a);
  level_scan_kernel<<<1, SCAN_THREADS, 0, stream>>>(a.tile_sel, tiles, n_selected, a.child_count, n_child_slots,
                                                    next_node_start, n_nodes_next);
  level_scatter_kernel<<<tiles, SWP_THREADS, 0, stream>>>(a);
}

// =============================================================================================
// K7  GRID_CENTER / JITTERED: segmented first-arg-min of the squared distance to a per-cell target
// =============================================================================================
//   GridCenterSampling::sample_points   tiling/Sampling.h:314-416
//   JitteredSampling::sample_points     tiling/Sampling.h:598-759
// Cells are runs of equal (key >> cell_shift) inside a node.  Per cell the reference selects
// std::min_element of squaredDistanceTo(target) = the FIRST minimum in Morton order.  Here every
// point evaluates its distance, a segmented inclusive min-scan (ties keep the earlier point) runs
// over the list in one pass, and the last point of each cell marks the winner in sel[].  Tiles are
// independent: a cell that begins in an earlier tile leaves its partial minimum in a descriptor and
// argmin_carry_kernel (one thread per tile) combines it with the end-of-tile minima of the tiles
// before.  (Until the second half of round 2 a single-thread look-back inside the kernel did this;
// ncu showed a quarter of all stall samples at the block barrier behind it, 33 polls per tile on
// average: 12.47 -> 11.27 ms per C3-shaped sweep of 100 M points with the wait taken out.)
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

// per-node quantities of JitteredSampling::sample_points, Sampling.h:621-660, from the node's bounds
__device__ __forceinline__ u32
jitter_node_setup(const double mn[3], const double mx[3], const SwArgminArgs& a, JitterNode& jn)
{
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

// The three permutation rows a launch needs (they depend on the node LEVEL only: start_index = 3 * (level + 1)
// mod 16, Sampling.h:684-693) for the three table sizes, staged in shared memory: lanes index them with
// different cells, which a __constant__ table would serialise.
#define JIT_ROW 64
#define JIT_TABLE_WORDS (3 * 3 * JIT_ROW) /* [size class 16 / 32 / 64][row][entry] */

// All 16 possible start indices, laid out the way the kernel wants them; filled once per device by
// jitter_build_tables_kernel.  (Staging straight from the __constant__ tables cost 14 % of select_argmin_kernel's stall
// samples: 576 constant loads per block, every lane at another address, plus the index divisions.)
__device__ u32 g_jit_staged[16][JIT_TABLE_WORDS];

__global__ void __launch_bounds__(256)
jitter_build_tables_kernel()
{
  const u32 start_index = blockIdx.x;
  for (u32 e = threadIdx.x; e < JIT_TABLE_WORDS; e += blockDim.x) {
    const u32 cls = e / (3 * JIT_ROW), row = (e / JIT_ROW) % 3, i = e % JIT_ROW;
    const u32 t = (start_index + row) % 16u;
    u32 v = 0;
    if (cls == 0)
      v = i < 16 ? PERMUTATIONS_16[t][i] : 0u;
    else if (cls == 1)
      v = i < 32 ? PERMUTATIONS_32[t][i] : 0u;
    else
      v = PERMUTATIONS_64[t][i];
    g_jit_staged[start_index][e] = v;
  }
}

__device__ __forceinline__ void
jitter_stage_tables(int node_level, u32* s_perm)
{
  const u32* __restrict__ src = g_jit_staged[(3u * (u32)(node_level + 1)) % 16u];
  for (u32 e = threadIdx.x; e < JIT_TABLE_WORDS; e += blockDim.x)
    s_perm[e] = src[e];
}

__device__ __forceinline__ void
jitter_target(u64 key, const JitterNode& jn, const u32* __restrict__ s_perm, double t[3])
{
  const u64 rel = key >> jn.shift;
  const u64 grid_mask = (1ull << (3 * jn.levels)) - 1ull;
  const u64 idx = rel & grid_mask;
  const u32 lmask = (1u << jn.levels) - 1u;
  u32 gx, gy, gz; // OctreeNodeIndex::to_grid_index, OctreeNodeIndex.h:357-363
  if (jn.levels <= 10) { // the usual grids (128 cells per axis: 21 bits): 32-bit arithmetic
    const u32 i32 = (u32)idx;
    gz = contract_bits_by_3_u32(i32) & lmask;
    gy = contract_bits_by_3_u32(i32 >> 1) & lmask;
    gx = contract_bits_by_3_u32(i32 >> 2) & lmask;
  } else {
    gz = (u32)contract_bits_by_3(idx) & lmask;
    gy = (u32)contract_bits_by_3(idx >> 1) & lmask;
    gx = (u32)contract_bits_by_3(idx >> 2) & lmask;
  }
  // len = min(cells, 64) is a power of two (cells = get_prev_power_of_two): x % len == x & (len - 1)
  const u32 len = jn.cells < 64u ? jn.cells : 64u;
  const u32 ix = (gy + gz) & (len - 1u), iy = (gx + gz) & (len - 1u), iz = (gx + gy) & (len - 1u);
  const u32* tab = s_perm + (jn.cells <= 16u ? 0u : (jn.cells <= 32u ? 1u : 2u)) * (3 * JIT_ROW);
  const u32 px = tab[ix] - 1u;
  const u32 py = tab[JIT_ROW + iy] - 1u;
  const u32 pz = tab[2 * JIT_ROW + iz] - 1u;
  // target = node_min + (g * cell + p * sub), Sampling.h:735-739 (mul, mul, add, add: no FMA)
  t[0] = jn.node_min[0] + ((double)gx * jn.grid_cell_size + (double)px * jn.permutation_cell_size);
  t[1] = jn.node_min[1] + ((double)gy * jn.grid_cell_size + (double)py * jn.permutation_cell_size);
  t[2] = jn.node_min[2] + ((double)gz * jn.grid_cell_size + (double)pz * jn.permutation_cell_size);
}

// Everything select_argmin_kernel needs per NODE, computed once per node instead of once per point: the
// node's bounds (the halving recurrence down to the node) and, for JITTERED, the grid quantities of
// Sampling.h:621-660 including the two exceptions it can throw (only for nodes that are really sampled).
__global__ void __launch_bounds__(256)
argmin_nodes_kernel(SwArgminArgs a)
{
  const u32 r = blockIdx.x * 256 + threadIdx.x;
  if (r >= a.n_nodes)
    return;
  const u64 key = a.in_key[a.node_start[r]] & SW_KEY_MASK;
  SwArgminNode nd;
  bounds_from_key(key, a.node_level + 1, a.bounds, nd.mn, nd.mx);
  nd.shift = a.cell_shift;
  nd.levels = 0;
  nd.cells = 1;
  nd.grid_cell_size = 0.0;
  nd.permutation_cell_size = 0.0;
  if (a.sampling == SW_JITTERED) {
    JitterNode jn;
    const u32 e = jitter_node_setup(nd.mn, nd.mx, a, jn);
    nd.shift = jn.shift;
    nd.levels = jn.levels;
    nd.cells = jn.cells;
    nd.grid_cell_size = jn.grid_cell_size;
    nd.permutation_cell_size = jn.permutation_cell_size;
    if (e) {
      bool active = true;
      if (a.allow_take_all)
        active = (u64)node_point_count(a.node_start, a.node_gcount, r) > a.max_points_per_node;
      if (active)
        atomicMax(a.error_flag, e);
    }
  }
  a.nodes[r] = nd;
}

// tile descriptors of the segmented scan: flag word + two 16-byte payload slots per tile
struct ArgminDesc
{
  double d;
  u32 pos;
  u32 has_head;
};

// Layout: keys, node ranks and head / tail flags are worked out in the BLOCKED layout (thread t owns elements
// [8t, 8t + 8) of the tile: runs are thread-local bit operations); the distances are evaluated in the STRIPED
// layout (element j * 256 + t: coalesced id / position loads) and handed over through shared memory; the
// segmented min-scan then runs thread-locally over 8 elements, across the 32 thread aggregates of a warp with
// 5 shuffle steps, across the warps through shared memory and across tiles by argmin_carry_kernel.
#ifndef ARGMIN_P1_UNROLL
#define ARGMIN_P1_UNROLL 4 /* elements of the distance phase whose loads are in flight together; fully unrolled (8) the kernel has 7 400 SASS lines and 14 % of its stall samples wait for instructions: 11.00 -> 10.60 ms per C3-shaped sweep */
#endif
#ifndef ARGMIN_MIN_CTAS
#define ARGMIN_MIN_CTAS 6 /* 40 registers.  4 -> 5 blocks per SM: 13.11 -> 12.48 ms per C3-shaped sweep of 100 M points; 5 -> 6 gave nothing while the kernel waited on the look-back, 10.30 -> 10.17 ms without it */
#endif
__global__ void __launch_bounds__(SWP_THREADS, ARGMIN_MIN_CTAS)
select_argmin_kernel(SwArgminArgs a, u64* __restrict__ status, u32* __restrict__ ticket)
{
  __shared__ u64 s_k[BLK_SLOTS]; // keys, then the distances (double bits) of the same elements
  __shared__ u32 s_r[BLK_SLOTS]; // node rank of every element, bit 31 = its node is not sampled
  __shared__ u64 s_prev, s_next;
  __shared__ u32 s_w[SWP_WARPS];
  __shared__ ArgminDesc s_wagg[SWP_WARPS];
  __shared__ u32 s_perm[JIT_TABLE_WORDS];
  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (a.sampling == SW_JITTERED)
    jitter_stage_tables(a.node_level, s_perm); // visible after the barriers of phase 0
  const u32 tile = blockIdx.x; // tiles are independent: what crosses a tile boundary is settled by argmin_carry_kernel
  const u64 base = (u64)tile * SW_SWEEP_TILE;
  const u64 e0 = base + 8ull * tid;
  const size_t n_tiles = (size_t)((a.count + SW_SWEEP_TILE - 1) / SW_SWEEP_TILE);
  ArgminDesc* first = reinterpret_cast<ArgminDesc*>(status);              // n_tiles x 16 B, zeroed by the launcher
  ArgminDesc* agg = reinterpret_cast<ArgminDesc*>(status + 2 * n_tiles);  // n_tiles x 16 B

  // ---- phase 0 (blocked): node ranks, which nodes are sampled, cell heads and tails --------------------------
  u32 hbits = 0, tbits = 0, abits = 0; // per element: starts a cell, ends a cell, its node is sampled
  {
    u64 k[BLK_ITEMS];
    u64 prev;
    if (tid == 0)
      s_next = (base + SW_SWEEP_TILE < a.count) ? (a.in_key[base + SW_SWEEP_TILE] & SW_KEY_MASK) : ~0ull;
    load_keys_blocked(a.in_key, base, a.count, s_k, &s_prev, k, prev); // keys past the end read as ~0
    const u64 next = (tid + 1 < SWP_THREADS) ? s_k[9 * (tid + 1)] : s_next;
    const u32 nvalid = e0 >= a.count ? 0u : (a.count - e0 < 8 ? (u32)(a.count - e0) : 8u);
    const u32 valid = (1u << nvalid) - 1u;
    u32 nh, unused;
    head_bits2(k, prev, a.node_shift, a.node_shift, e0 == 0, nh, unused);
    nh &= valid;
    u32 hexcl;
    block_scan_packed(__popc(nh), s_w, hexcl);
    u32 rank = a.tile_rank0[tile] + hexcl - 1u; // node of the element before my first one
    bool active = true;
    int es = a.node_shift; // effective shift of the selection cells of the current node
#pragma unroll
    for (int j = 0; j < BLK_ITEMS; ++j) {
      const bool is_valid = (valid >> j) & 1u;
      const bool node_head = (nh >> j) & 1u;
      if (node_head)
        ++rank;
      if (is_valid && (node_head || j == 0)) {
        active = true;
        if (a.allow_take_all)
          active = (u64)node_point_count(a.node_start, a.node_gcount, rank) > a.max_points_per_node;
        const int cs = (a.sampling == SW_JITTERED) ? a.nodes[rank].shift : a.cell_shift;
        es = cs < a.node_shift ? cs : a.node_shift;
      }
      const u64 x_prev = k[j] ^ (j ? k[j - 1] : prev);
      const u64 x_next = k[j] ^ (j + 1 < BLK_ITEMS ? k[j + 1] : next);
      const bool first = (e0 + j == 0);
      const bool head = !active || node_head || first || ((x_prev >> es) != 0ull);
      const bool tail = !active || ((x_next >> es) != 0ull); // keys past the end differ in bit 63
      if (is_valid) {
        hbits |= (head ? 1u : 0u) << j;
        tbits |= (tail ? 1u : 0u) << j;
        abits |= (active ? 1u : 0u) << j;
      } else {
        hbits |= 1u << j; // padding behaves like a run of its own
      }
      s_r[BLK_PAD(8 * tid + j)] = rank | (active ? 0u : 0x80000000u);
    }
  }
  __syncthreads();

  // ---- phase 1 (striped): squared distance of every sampled point to the target of its cell ------------------
  constexpr int P1_UNROLL = ARGMIN_P1_UNROLL;
#pragma unroll P1_UNROLL
  for (int j = 0; j < SWP_ITEMS; ++j) {
    const u32 p = j * SWP_THREADS + tid;
    const u64 i = base + p;
    double d = 0.0;
    if (i < a.count) {
      const u32 r = s_r[BLK_PAD(p)];
      if (!(r & 0x80000000u)) {
        const u64 k = s_k[BLK_PAD(p)];
        const u32 idx = a.in_idx ? a.in_idx[i] : (u32)i;
        const double pt[3] = { a.pos_sorted[3 * (u64)idx], a.pos_sorted[3 * (u64)idx + 1],
                               a.pos_sorted[3 * (u64)idx + 2] };
        double t[3];
        const SwArgminNode* nd = a.nodes + r;
        if (a.sampling == SW_JITTERED) {
          JitterNode jn;
          jn.shift = nd->shift;
          jn.levels = nd->levels;
          jn.cells = nd->cells;
          jn.node_min[0] = nd->mn[0];
          jn.node_min[1] = nd->mn[1];
          jn.node_min[2] = nd->mn[2];
          jn.grid_cell_size = nd->grid_cell_size;
          jn.permutation_cell_size = nd->permutation_cell_size;
          jitter_target(k, jn, s_perm, t);
        } else {
          double mn[3], mx[3];
          if (a.cand_level >= a.node_level) { // the usual case: continue from the node's bounds
            mn[0] = nd->mn[0];
            mn[1] = nd->mn[1];
            mn[2] = nd->mn[2];
            mx[0] = nd->mx[0];
            mx[1] = nd->mx[1];
            mx[2] = nd->mx[2];
            bounds_continue(k, a.node_level + 1, a.cand_level + 1, mn, mx);
          } else { // spacing coarser than the node: the candidate cell is an ancestor of the node
            bounds_from_key(k, a.cand_level + 1, a.bounds, mn, mx);
          }
          // AABB::getCenter = min + extent()/2 (math/AABB.h:70)
          t[0] = mn[0] + (mx[0] - mn[0]) * 0.5;
          t[1] = mn[1] + (mx[1] - mn[1]) * 0.5;
          t[2] = mn[2] + (mx[2] - mn[2]) * 0.5;
        }
        d = squared_distance(pt, t);
      }
    }
    s_k[BLK_PAD(p)] = (u64)__double_as_longlong(d); // the slot's key is not needed any more
  }
  __syncthreads();

  // ---- phase 2 (blocked): segmented first-arg-min ----------------------------------------------------------------
  double d[BLK_ITEMS];
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j)
    d[j] = __longlong_as_double((long long)s_k[9 * tid + j]);
  // thread aggregate: running minimum of the run that is open at the end of my 8 elements
  ArgminVal run;
  run.d = d[0];
  run.pos = (u32)e0;
#pragma unroll
  for (int j = 1; j < BLK_ITEMS; ++j) {
    ArgminVal v;
    v.d = d[j];
    v.pos = (u32)(e0 + j);
    run = ((hbits >> j) & 1u) ? v : argmin_op(run, v);
  }
  const bool my_head = hbits != 0;
  // inclusive segmented scan of the thread aggregates inside the warp
  ArgminVal inc = run;
  bool finc = my_head;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    ArgminVal up;
    up.d = __shfl_up_sync(0xffffffffu, inc.d, o);
    up.pos = __shfl_up_sync(0xffffffffu, inc.pos, o);
    const bool fup = __shfl_up_sync(0xffffffffu, (int)finc, o) != 0;
    if (lane >= (u32)o) {
      if (!finc)
        inc = argmin_op(up, inc);
      finc = finc || fup;
    }
  }
  // what reaches my first element from the earlier lanes of the warp
  ArgminVal lane_in;
  lane_in.d = __shfl_up_sync(0xffffffffu, inc.d, 1);
  lane_in.pos = __shfl_up_sync(0xffffffffu, inc.pos, 1);
  const bool lane_in_cut = __shfl_up_sync(0xffffffffu, (int)finc, 1) != 0; // a head in the earlier lanes
  if (lane == 31) {
    s_wagg[warp].d = inc.d;
    s_wagg[warp].pos = inc.pos;
    s_wagg[warp].has_head = finc ? 1u : 0u;
  }
  __syncthreads();

  // ---- tile aggregate: the run that is open at the end of the tile, for the tiles behind this one --------------
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
    agg[tile] = mine; // the run that is open at the end of the tile (the whole tile when it has no head)
  }

  // ---- carry into my first element = (tile carry, previous warps, previous lanes), then the winners ------------
  // (no barrier needed here: s_wagg was complete at the barrier above, thread 0 only read it)
  ArgminVal cin;
  cin.d = 0.0;
  cin.pos = 0;
  bool cin_valid = false;
  bool from_tile_start = true; // no cell head between the start of the tile and my first element
  for (u32 w = 0; w < warp; ++w) {
    ArgminVal wv;
    wv.d = s_wagg[w].d;
    wv.pos = s_wagg[w].pos;
    if (s_wagg[w].has_head || !cin_valid)
      cin = wv;
    else
      cin = argmin_op(cin, wv);
    cin_valid = true;
    from_tile_start = from_tile_start && !s_wagg[w].has_head;
  }
  if (lane > 0) {
    if (lane_in_cut || !cin_valid)
      cin = lane_in;
    else
      cin = argmin_op(cin, lane_in);
    cin_valid = true;
    from_tile_start = from_tile_start && !lane_in_cut;
  }
  const u32 nvalid = e0 >= a.count ? 0u : (a.count - e0 < 8 ? (u32)(a.count - e0) : 8u);
  bool open = cin_valid; // `cur` continues a run that started before my first element
  ArgminVal cur = cin;
#pragma unroll
  for (int j = 0; j < BLK_ITEMS; ++j) {
    ArgminVal v;
    v.d = d[j];
    v.pos = (u32)(e0 + j);
    if ((hbits >> j) & 1u)
      from_tile_start = false;
    if (((hbits >> j) & 1u) || !open)
      cur = v;
    else
      cur = argmin_op(cur, v);
    open = true;
    if ((u32)j < nvalid && ((tbits >> j) & 1u) && ((abits >> j) & 1u)) {
      if (from_tile_start && tile != 0) {
        // the cell may have begun in an earlier tile: its winner is decided by argmin_carry_kernel
        ArgminDesc f;
        f.d = cur.d;
        f.pos = cur.pos;
        f.has_head = 1u; // "valid"
        first[tile] = f;
      } else {
        a.sel[cur.pos] = 1;
      }
    }
  }
}

// One thread per tile whose first cell ends in the tile: the minimum of the part of the cell that lies in earlier
// tiles (their end-of-tile aggregates, back to the first tile that holds a cell head) against the part in this tile;
// ties keep the earlier point.
__global__ void __launch_bounds__(256)
argmin_carry_kernel(const u64* __restrict__ status, u32 n_tiles, unsigned char* __restrict__ sel)
{
  const u32 tile = blockIdx.x * 256 + threadIdx.x;
  if (tile >= n_tiles)
    return;
  const ArgminDesc* first = reinterpret_cast<const ArgminDesc*>(status);
  const ArgminDesc* agg = reinterpret_cast<const ArgminDesc*>(status + 2 * (size_t)n_tiles);
  const ArgminDesc f = first[tile];
  if (!f.has_head)
    return;
  ArgminVal best;
  best.d = f.d;
  best.pos = f.pos;
  for (long long t = (long long)tile - 1; t >= 0; --t) {
    const ArgminDesc e = agg[t];
    ArgminVal ev;
    ev.d = e.d;
    ev.pos = e.pos;
    best = argmin_op(ev, best); // the earlier tile's points come first
    if (e.has_head)
      break;
  }
  sel[best.pos] = 1;
}

void
launch_select_argmin(const SwArgminArgs& a, u64* status, u32* ticket, cudaStream_t stream)
{
  const size_t tiles = sweep_tiles(a.count);
  (void)ticket;
  if (a.sampling == SW_JITTERED) { // the staged permutation tables exist once per device
    static bool built[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !built[dev]) {
      // once per device and process: wait for it, handles on other streams of this device rely on the flag
      jitter_build_tables_kernel<<<16, 256, 0, stream>>>();
      cudaStreamSynchronize(stream);
      if (dev >= 0 && dev < 64)
        built[dev] = true;
    }
  }
  cudaMemsetAsync(status, 0, tiles * sizeof(ArgminDesc), stream); // the "first cell of the tile" descriptors
  argmin_nodes_kernel<<<(a.n_nodes + 255) / 256, 256, 0, stream>>>(a);
  select_argmin_kernel<<<(u32)tiles, SWP_THREADS, 0, stream>>>(a, status, ticket);
  argmin_carry_kernel<<<(u32)((tiles + 255) / 256), 256, 0, stream>>>(status, (u32)tiles, a.sel);
}

// =============================================================================================
// K9  MIN_DISTANCE: greedy minimum-distance selection in Morton order
// =============================================================================================
//   PoissonDiskSampling::sample_points   tiling/Sampling.h:421-471
//   SparseGrid::add / GridCell::isDistant datastructures/SparseGrid.cpp:116-146, GridCell.cpp:41-58
// Reference semantics: walk the node's points in Morton order; accept a point iff no previously
// ACCEPTED point lies closer than the spacing (squared distance < (double)(float)(s*s), strict).
// The SparseGrid only accelerates the search (its 27-cell stencil covers every point within the
// spacing because its cells are 5 spacings wide), so the result is the lexicographically first
// maximal independent set of the "closer than spacing" graph.
//
// Parallel formulation (exact): only ACCEPTED points ever influence a decision, and a point can only
// be influenced by points that precede it in Morton order.  The list is sorted, so a Morton cell whose
// side is >= spacing is a contiguous run, every point within the spacing lies in the same or an
// adjacent cell, and all points of an adjacent cell with a smaller Morton code precede all points of
// this cell.  Cells are therefore processed as a wavefront in Morton order: a cell waits until its (at
// most 26) earlier neighbour cells of the same node have published their accepted points, then walks
// its own points in order, 32 (or 256) at a time: every point is first tested against the accepted
// points known so far (neighbours + own), the survivors of a batch are resolved in order by one warp.
// Cells are handed out through an atomic ticket in Morton order, so every cell a group waits for is
// already owned by a resident group (the decoupled look-back argument): no deadlock, no grid barrier,
// and every point is touched once.  The critical path is the longest chain of adjacent cells with
// increasing Morton code inside one node (measured: ~4 000 cells for a 128 x 128 x 4 terrain slab).
// md_wave_warp_kernel gives a cell to one warp (sparse levels: few points per cell, many cells),
// md_wave_cta_kernel to a whole CTA (dense levels: thousands of points per cell on the critical path).
//
// (Round 1 used fix-point rounds over an "undecided" work list in a cooperative kernel: exact too, but
// the blocked-on chains run through points that are about to be rejected -- 13 000 rounds for one
// million terrain points, 1.2 s.)
#define MD_UNDECIDED 0
#define MD_ACCEPTED 1
#define MD_REJECTED 2
#define MD_NBR_SLOTS 27 /* slot 0 = count, then up to 26 earlier neighbour cells */
#define MD_OWN_CAP 64   /* accepted points of one cell: pairwise >= spacing apart inside a cube of side < 2 spacings */
#define MD_DESC_DONE (1ull << 63) /* cell descriptor: done | accepted count << 32 | offset into acc_xyz */

static cudaError_t
grow(SwGrowBuf& b, size_t bytes)
{
  if (bytes <= b.cap)
    return cudaSuccess;
  if (b.p)
    cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  const size_t want = bytes + bytes / 4;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e == cudaSuccess)
    b.cap = want;
  return e;
}

void
free_min_distance_scratch(SwMinDistScratch& sc)
{
  SwGrowBuf* all[] = { &sc.hmask, &sc.cell_start, &sc.cell_tile_rank0, &sc.state, &sc.lpos, &sc.desc,  &sc.acc_xyz,    &sc.hkeys,
                       &sc.hvals,      &sc.nbr,             &sc.deps,  &sc.queue, &sc.cell_active, &sc.counters };
  for (SwGrowBuf* b : all) {
    if (b->p)
      cudaFree(b->p);
    b->p = nullptr;
    b->cap = 0;
  }
}

// per point: activity (take-all nodes are never sampled), compact position copy; counts the active points
__global__ void __launch_bounds__(SWP_THREADS)
md_setup_kernel(SwMinDistArgs a, unsigned char* __restrict__ state, double* __restrict__ lpos,
                u32* __restrict__ n_active)
{
  __shared__ u32 s_w[SWP_WARPS];
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 tile = blockIdx.x;
  const u64 base = (u64)tile * SW_SWEEP_TILE;
  const u32 lt = lanemask_lt();
  u32 nmask[SWP_ITEMS];
  u32 wn = 0;
#pragma unroll
  for (int j = 0; j < SWP_ITEMS; ++j) {
    const u64 i = base + item_pos(warp, lane, j);
    bool nh = false;
    if (i < a.count) {
      const u64 k = a.in_key[i] & SW_KEY_MASK;
      nh = (i == 0) || (k >> a.node_shift) != ((a.in_key[i - 1] & SW_KEY_MASK) >> a.node_shift);
    }
    nmask[j] = __ballot_sync(0xffffffffu, nh);
    wn += __popc(nmask[j]);
  }
  u32 tn;
  const u32 nexcl = warp_totals_exclusive(wn, warp, lane, s_w, tn);
  u32 nrun = a.tile_rank0[tile] + nexcl;
  u32 my_active = 0;
#pragma unroll
  for (int j = 0; j < SWP_ITEMS; ++j) {
    const u64 i = base + item_pos(warp, lane, j);
    const u32 self = lt | (1u << lane);
    const u32 node_rank = nrun + __popc(nmask[j] & self) - 1;
    nrun += __popc(nmask[j]);
    if (i < a.count) {
      bool active = true;
      if (a.allow_take_all) {
        const u32 cnt = node_point_count(a.node_start, a.node_gcount, node_rank);
        active = (u64)cnt > a.max_points_per_node;
      }
      // AdaptivePoissonDiskSampling: the counter starts at nth - 1, so positions 0, nth, 2 nth, ... of
      // the node's range are analysed, every other point is rejected without touching the grid
      if (a.nth_point > 1 && ((u32)i - a.node_start[node_rank]) % a.nth_point != 0)
        active = false;
      state[i] = active ? MD_UNDECIDED : MD_REJECTED;
      my_active += active ? 1u : 0u;
      if (lpos) {
        const u64 idx = a.in_idx[i];
        lpos[3 * i] = a.pos_sorted[3 * idx];
        lpos[3 * i + 1] = a.pos_sorted[3 * idx + 1];
        lpos[3 * i + 2] = a.pos_sorted[3 * idx + 2];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    my_active += __shfl_xor_sync(0xffffffffu, my_active, o);
  if (lane == 0 && my_active)
    atomicAdd(n_active, my_active);
}

// Uniform hash on purpose.  A locality-preserving variant (hash the 4 x 4 x 4 block, keep the cell's place inside
// a 64-slot group) halves md_hash_insert_kernel but makes md_neighbors_kernel slower (109 vs 91 ms over the levels
// of 125 M terrain points): two thirds of the 26 neighbour probes look for cells that do not exist and have to
// walk the clustered groups to the next empty slot.
__device__ __forceinline__ u32
md_hash(u64 code, u32 mask)
{
  code ^= code >> 33;
  code *= 0xff51afd7ed558ccdull;
  code ^= code >> 33;
  return (u32)code & mask;
}

// The occupied cells are found through a hash table over 4 x 4 x 4 BLOCKS of cells (block = cell code >> 6): one
// entry holds the block's first cell (cells are in Morton order, so a block's cells are consecutive) and a 64-bit
// occupancy mask; cell = first + popc(mask below the cell's bit).  A cell's 26 neighbours lie in at most 8 blocks,
// neighbours that do not exist are answered by the mask instead of a probe sequence, and the table has a fraction
// of the entries of a per-cell table (round 2: 26 random probes per cell were 45 ms of a 237 ms step).
// One thread per cell: is there anything to analyse in the cell (cells of take-all nodes are not); the first cell
// of every block inserts the block.
__global__ void __launch_bounds__(256)
md_hash_insert_kernel(const u64* __restrict__ in_key, const u32* __restrict__ cell_start, u32 n_cells, int cell_shift,
                      const unsigned char* __restrict__ state, unsigned char* __restrict__ cell_active,
                      u64* __restrict__ hkeys, u32* __restrict__ hvals, u64* __restrict__ hmask, u32 mask)
{
  const u32 c = blockIdx.x * 256 + threadIdx.x;
  if (c >= n_cells)
    return;
  const u32 b = cell_start[c], e = cell_start[c + 1];
  bool active = false;
  for (u32 i = b; i < e && !active; ++i) // analysed cells answer at their first point
    active = state[i] == MD_UNDECIDED;
  cell_active[c] = active ? 1 : 0;
  const u64 code = (in_key[b] & SW_KEY_MASK) >> cell_shift;
  const u64 block = code >> 6;
  if (c > 0 && (((in_key[cell_start[c - 1]] & SW_KEY_MASK) >> cell_shift) >> 6) == block)
    return; // not the first cell of its block
  u64 occupied = 1ull << (code & 63ull);
  for (u32 k = c + 1; k < n_cells; ++k) {
    const u64 other = (in_key[cell_start[k]] & SW_KEY_MASK) >> cell_shift;
    if ((other >> 6) != block)
      break;
    occupied |= 1ull << (other & 63ull);
  }
  u32 slot = md_hash(block, mask);
  while (true) {
    const u64 prev = atomicCAS(&hkeys[slot], 0ull, block + 1);
    if (prev == 0ull) {
      hvals[slot] = c;
      hmask[slot] = occupied;
      return;
    }
    slot = (slot + 1) & mask;
  }
}

// Neighbour cells of every analysed cell, same node only: the EARLIER ones (smaller Morton code: their
// accepted points constrain this cell) and the LATER ones (they wait for this cell).  Cells without
// analysed points never hold an accepted point and are left out on both sides.  A cell whose earlier
// list is empty is ready at once and goes straight into the ready queue.
// nbr row: [0] = earlier count | later count << 8, [1 ...] = earlier cells, then later cells.
// counters: [0] pop ticket, [1] next free slot of acc_xyz, [2] error flag, [3] active points,
//           [4] cell count (node_rle), [5] analysed cells, [6] push ticket
// The ready queue is split into mq.n independent queues (cell c belongs to queue (c / 64) mod n), each with its
// own ticket, push and analysed-cell counters on separate 128-byte lines: at the deep levels a sweep level has
// tens of millions of cells, and one pop and one push per cell on a single address serialise in one L2 slice
// (measured: 204 -> 75 ms for 54 M cells).  Queue q owns queue[q * cap, (q + 1) * cap).
#define MD_QUEUE_CHUNK 64
#define MD_MAX_QUEUES 64
#define MDQ_POP 0
#define MDQ_PUSH 32
#define MDQ_ANALYSED 64
#define MDQ_STRIDE 96
#define MDQ_BASE 32 /* u32 index of queue 0's counters inside `counters` */
struct MdQueues
{
  u32 n;   // number of queues (1 .. MD_MAX_QUEUES)
  u32 cap; // entries per queue
};

__device__ __forceinline__ u32
md_queue_of(u32 cell, const MdQueues& mq)
{
  return (cell / MD_QUEUE_CHUNK) % mq.n;
}

__device__ __forceinline__ u32*
md_queue_counters(u32* counters, u32 q)
{
  return counters + MDQ_BASE + q * MDQ_STRIDE;
}

// neighbour row + dependency counter of one analysed cell; returns the number of earlier neighbours
__device__ __forceinline__ u32
md_neighbors_of(u32 c, const u64* __restrict__ in_key, const u32* __restrict__ cell_start, int cell_shift,
                int cell_levels, int node_levels, const unsigned char* __restrict__ cell_active,
                const u64* __restrict__ hkeys, const u32* __restrict__ hvals, const u64* __restrict__ hmask, u32 mask,
                u32* __restrict__ nbr, u32* __restrict__ deps)
{
  u32* out = nbr + (size_t)c * MD_NBR_SLOTS;
  u32 later[26];
  u32 n_early = 0, n_late = 0;
  const int below = cell_levels - node_levels; // cell levels below the node
  if (below > 0) {
    const u64 code = (in_key[cell_start[c]] & SW_KEY_MASK) >> cell_shift;
    const long long side = 1ll << cell_levels;
    const long long x = (long long)contract_bits_by_3(code >> 2);
    const long long y = (long long)contract_bits_by_3(code >> 1);
    const long long z = (long long)contract_bits_by_3(code);
    const u64 node_prefix = code >> (3 * below);
    // the (at most 8) blocks the neighbourhood touches, looked up once each: slot = which side of the own block
    // per axis (0 = own block's coordinate, 1 = the adjacent one)
    u64 blk_mask[8];
    u32 blk_first[8];
    u32 blk_known = 0;
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz) {
          if (!dx && !dy && !dz)
            continue;
          const long long X = x + dx, Y = y + dy, Z = z + dz;
          if (X < 0 || Y < 0 || Z < 0 || X >= side || Y >= side || Z >= side)
            continue;
          const u64 nc = expand_bits_by_3((u64)Z) | (expand_bits_by_3((u64)Y) << 1) | (expand_bits_by_3((u64)X) << 2);
          if ((nc >> (3 * below)) != node_prefix)
            continue;
          const u32 bi = (((X >> 2) != (x >> 2)) ? 4u : 0u) | (((Y >> 2) != (y >> 2)) ? 2u : 0u) |
                         (((Z >> 2) != (z >> 2)) ? 1u : 0u);
          if (!((blk_known >> bi) & 1u)) {
            const u64 block = nc >> 6;
            u64 m = 0ull;
            u32 f = 0u;
            u32 slot = md_hash(block, mask);
            while (true) {
              const u64 k = hkeys[slot];
              if (k == 0ull)
                break;
              if (k == block + 1) {
                f = hvals[slot];
                m = hmask[slot];
                break;
              }
              slot = (slot + 1) & mask;
            }
            blk_mask[bi] = m;
            blk_first[bi] = f;
            blk_known |= 1u << bi;
          }
          const u32 bit = (u32)(nc & 63ull);
          const u64 m = blk_mask[bi];
          if (!((m >> bit) & 1ull))
            continue; // no such cell
          const u32 other = blk_first[bi] + (u32)__popcll(m & ((1ull << bit) - 1ull));
          if (cell_active[other]) {
            if (nc < code)
              out[1 + n_early++] = other;
            else
              later[n_late++] = other;
          }
        }
  }
  for (u32 k = 0; k < n_late; ++k)
    out[1 + n_early + k] = later[k];
  out[0] = n_early | (n_late << 8);
  deps[c] = n_early;
  return n_early;
}

__global__ void __launch_bounds__(256)
md_neighbors_kernel(const u64* __restrict__ in_key, const u32* __restrict__ cell_start, u32 n_cells, int cell_shift,
                    int cell_levels, int node_levels, const unsigned char* __restrict__ cell_active,
                    const u64* __restrict__ hkeys, const u32* __restrict__ hvals, const u64* __restrict__ hmask, u32 mask,
                    u32* __restrict__ nbr, u32* __restrict__ deps, u32* __restrict__ queue, u32* __restrict__ counters,
                    MdQueues mq)
{
  const u32 c = blockIdx.x * 256 + threadIdx.x;
  const bool act = c < n_cells && cell_active[c];
  u32 n_early_out = 1;
  if (act)
    n_early_out = md_neighbors_of(c, in_key, cell_start, cell_shift, cell_levels, node_levels, cell_active, hkeys, hvals,
                                  hmask, mask, nbr, deps);
  // the 32 cells of a warp belong to one queue (MD_QUEUE_CHUNK = 64 consecutive cells): one atomic per warp for
  // the analysed-cell count and one for the cells that are ready at once
  const u32 lane = threadIdx.x & 31;
  const u32 q = md_queue_of(c, mq);
  u32* qc = md_queue_counters(counters, q);
  const u32 am = __ballot_sync(0xffffffffu, act);
  const u32 rm = __ballot_sync(0xffffffffu, act && n_early_out == 0);
  u32 base = 0;
  if (lane == 0) {
    if (am)
      atomicAdd(qc + MDQ_ANALYSED, (u32)__popc(am));
    if (rm)
      base = atomicAdd(qc + MDQ_PUSH, (u32)__popc(rm));
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  if (act && n_early_out == 0)
    queue[(size_t)q * mq.cap + base + __popc(rm & lanemask_lt())] = c;
}

#define MD_QUEUE_EMPTY 0xffffffffu

__device__ __forceinline__ u32
ld_acquire_u32(const u32* p)
{
  u32 v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void
st_release_u32(u32* p, u32 v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Takes the next ready cell of queue q.  Every analysed cell is pushed exactly once (when its last earlier
// neighbour finishes, or by md_neighbors_kernel when it has none), so tickets below the queue's analysed-cell
// count are always served; whoever waits, waits on its own queue slot.  Returns MD_QUEUE_EMPTY when all cells of
// the queue are handed out.  Called by one thread.
__device__ __forceinline__ u32
md_pop_cell(const u32* queue, u32* counters, u32 q, const MdQueues& mq)
{
  u32* qc = md_queue_counters(counters, q);
  const u32 t = atomicAdd(qc + MDQ_POP, 1u);
  if (t >= __ldcg(qc + MDQ_ANALYSED))
    return MD_QUEUE_EMPTY;
  u32 c;
  const u32* slot = queue + (size_t)q * mq.cap + t;
  while ((c = ld_acquire_u32(slot)) == MD_QUEUE_EMPTY)
    __nanosleep(100);
  return c;
}

// The same over all queues, starting at the caller's home queue `q` (updated): a group works on its home queue
// until that is handed out, then helps with the next one.  Every queue keeps its home groups until it is empty,
// so a pushed cell is always claimed.
__device__ __forceinline__ u32
md_pop_any(const u32* queue, u32* counters, u32& q, u32& visited, const MdQueues& mq)
{
  while (visited < mq.n) {
    const u32 c = md_pop_cell(queue, counters, q, mq);
    if (c != MD_QUEUE_EMPTY)
      return c;
    q = (q + 1) % mq.n;
    ++visited;
  }
  return MD_QUEUE_EMPTY;
}

// This cell is finished (its accepted points and descriptor are written and fenced): release the later
// neighbours, queue those that have become ready.  Called by up to 26 threads, one per later neighbour.
__device__ __forceinline__ void
md_release_later(u32 cell, u32* deps, u32* queue, u32* counters, const MdQueues& mq)
{
  if (atomicSub(&deps[cell], 1u) == 1u) {
    __threadfence(); // everything the other earlier neighbours published is visible before the push
    const u32 q = md_queue_of(cell, mq);
    st_release_u32(queue + (size_t)q * mq.cap + atomicAdd(md_queue_counters(counters, q) + MDQ_PUSH, 1u), cell);
  }
}

// squared distance exactly as Vector3::squaredDistanceTo (math/Vector3.h:55-62): x*x + y*y + z*z, no FMA
__device__ __forceinline__ bool
md_in_range(double px, double py, double pz, double qx, double qy, double qz, double threshold)
{
  const double dx = px - qx, dy = py - qy, dz = pz - qz;
  return dx * dx + dy * dy + dz * dz < threshold;
}

// Tests one point against a list of accepted points; the warp leaves the loop as soon as none of its
// lanes is a candidate any more (in a dense cell the first few accepted points reject almost everything).
__device__ __forceinline__ bool
md_filter(bool cand, double px, double py, double pz, const double* acc, u32 n, double threshold)
{
  for (u32 k = 0; k < n; ++k) {
    if ((k & 3u) == 0 && !__any_sync(0xffffffffu, cand))
      break;
    if (cand && md_in_range(px, py, pz, acc[3 * k], acc[3 * k + 1], acc[3 * k + 2], threshold))
      cand = false;
  }
  return cand;
}

// Resolves the surviving candidates of one warp-wide batch in lane (= point) order: the first candidate
// is accepted, every later candidate within the spacing of it is dropped, and so on.  The accepted
// positions are appended to `own` at [n_own ...); returns whether this lane's point was accepted.
__device__ __forceinline__ bool
md_resolve_warp(bool cand, double px, double py, double pz, double threshold, double* own, u32& n_own, u32* error_flag)
{
  const u32 lane = threadIdx.x & 31;
  bool accepted = false;
  u32 m = __ballot_sync(0xffffffffu, cand);
  while (m) {
    const int l = __ffs(m) - 1;
    const double ax = __shfl_sync(0xffffffffu, px, l);
    const double ay = __shfl_sync(0xffffffffu, py, l);
    const double az = __shfl_sync(0xffffffffu, pz, l);
    if ((int)lane == l) {
      accepted = true;
      cand = false;
      if (n_own < MD_OWN_CAP) {
        own[3 * n_own] = ax;
        own[3 * n_own + 1] = ay;
        own[3 * n_own + 2] = az;
      } else {
        atomicExch(error_flag, 1u); // cannot happen for cells of side < 2 spacings; reported, never silent
      }
    } else if (cand && md_in_range(px, py, pz, ax, ay, az, threshold)) {
      cand = false;
    }
    if (n_own < MD_OWN_CAP)
      ++n_own;
    m = __ballot_sync(0xffffffffu, cand);
  }
  return accepted;
}

// counters: [0] cell ticket, [1] next free slot of acc_xyz, [2] error flag, [3] active points,
//           [4] cell count (node_rle)
#define MDW_WARP_NBR_CAP 96
#define MDW_WARP_CAP (MD_OWN_CAP + MDW_WARP_NBR_CAP)

// one warp per cell (sparse levels)
__global__ void __launch_bounds__(256)
md_wave_warp_kernel(const double* __restrict__ P, const u32* __restrict__ cell_start, const u32* __restrict__ nbr,
                    unsigned char* __restrict__ state, u64* desc, double* acc_xyz, u32* deps, u32* queue,
                    u32* counters, double threshold, MdQueues mq)
{
  __shared__ double s_acc[8][MDW_WARP_CAP * 3];
  const u32 lane = threadIdx.x & 31;
  double* own = s_acc[threadIdx.x >> 5];  // own accepted points first: they reject most of a dense cell
  double* nacc = own + MD_OWN_CAP * 3;    // then the neighbours' accepted points
  u32 home = (blockIdx.x * 8u + (threadIdx.x >> 5)) % mq.n, visited = 0;
  while (true) {
    u32 c = 0;
    if (lane == 0)
      c = md_pop_any(queue, counters, home, visited, mq);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c == MD_QUEUE_EMPTY)
      break;
    const u32 b = cell_start[c], e = cell_start[c + 1];
    // every earlier neighbour is finished: fetch their accepted points (one lane per neighbour)
    const u32* nb = nbr + (size_t)c * MD_NBR_SLOTS;
    const u32 nn = nb[0] & 0xffu, n_late = nb[0] >> 8;
    u64 d = 0;
    if (lane < nn)
      d = __ldcg(desc + nb[1 + lane]);
    const u32 ncnt = (u32)(d >> 32) & 0xffffu;
    const u32 noff = (u32)d;
    u32 incl = ncnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o)
        incl += t;
    }
    const u32 total = __shfl_sync(0xffffffffu, incl, 31);
    const bool staged = total <= MDW_WARP_NBR_CAP;
    if (staged) {
      const u32 first = incl - ncnt;
      for (u32 k = 0; k < ncnt; ++k) {
        const double* src = acc_xyz + 3 * (size_t)(noff + k);
        nacc[3 * (first + k)] = __ldcg(src);
        nacc[3 * (first + k) + 1] = __ldcg(src + 1);
        nacc[3 * (first + k) + 2] = __ldcg(src + 2);
      }
    }
    __syncwarp();
    u32 n_own = 0;
    for (u32 base = b; base < e; base += 32) {
      const u32 i = base + lane;
      const bool valid = i < e && state[i] == MD_UNDECIDED;
      double px = 0, py = 0, pz = 0;
      if (valid) {
        px = P[3 * (size_t)i];
        py = P[3 * (size_t)i + 1];
        pz = P[3 * (size_t)i + 2];
      }
      bool cand = md_filter(valid, px, py, pz, own, n_own, threshold);
      if (staged) {
        cand = md_filter(cand, px, py, pz, nacc, total, threshold);
      } else { // more neighbour points than the staging area holds: stream them from L2
        for (u32 j = 0; j < nn; ++j) {
          const u32 o = __shfl_sync(0xffffffffu, noff, j), cn = __shfl_sync(0xffffffffu, ncnt, j);
          for (u32 k = 0; k < cn; ++k) {
            const double* src = acc_xyz + 3 * (size_t)(o + k);
            if (cand && md_in_range(px, py, pz, __ldcg(src), __ldcg(src + 1), __ldcg(src + 2), threshold))
              cand = false;
          }
        }
      }
      const bool accepted = md_resolve_warp(cand, px, py, pz, threshold, own, n_own, counters + 2);
      __syncwarp();
      if (valid)
        state[i] = accepted ? MD_ACCEPTED : MD_REJECTED;
    }
    // publish the accepted points of this cell in the cell's own range of acc_xyz (one slot per point of the
    // level, so no allocation is needed)
    const u32 off = b;
    if (n_own) {
      for (u32 k = lane; k < n_own; k += 32) {
        double* dst = acc_xyz + 3 * (size_t)(off + k);
        __stcg(dst, own[3 * k]);
        __stcg(dst + 1, own[3 * k + 1]);
        __stcg(dst + 2, own[3 * k + 2]);
      }
      __threadfence();
    }
    if (lane == 0)
      __stcg(desc + c, MD_DESC_DONE | ((u64)n_own << 32) | off);
    __threadfence();
    __syncwarp();
    if (lane < n_late)
      md_release_later(nb[1 + nn + lane], deps, queue, counters, mq);
  }
}

// one THREAD per cell (the deep, sparse levels: tens of millions of cells holding a handful of points each; a warp
// per cell leaves 30 lanes idle and is bound by the per-cell latency of queue, descriptor and neighbour loads, so
// 32 times as many cells in flight win).  The accepted points of the cell go straight into acc_xyz at a range
// reserved for the cell's point count (acc_xyz holds one slot per point of the level), which doubles as the list
// the cell's later points are tested against.
__global__ void __launch_bounds__(256)
md_wave_thread_kernel(const double* __restrict__ P, const u32* __restrict__ cell_start, const u32* __restrict__ nbr,
                      unsigned char* __restrict__ state, u64* desc, double* acc_xyz, u32* deps, u32* queue,
                      u32* counters, double threshold, MdQueues mq)
{
  const u32 lane = threadIdx.x & 31;
  u32 home = (blockIdx.x * 8u + (threadIdx.x >> 5)) % mq.n, visited = 0;
  while (visited < mq.n) {
    // one ticket atomic per 32 cells: tens of millions of same-address atomics serialise in one L2 slice
    __syncwarp();
    u32* qc = md_queue_counters(counters, home);
    u32 t0 = 0;
    if (lane == 0)
      t0 = atomicAdd(qc + MDQ_POP, 32u);
    t0 = __shfl_sync(0xffffffffu, t0, 0);
    const u32 analysed = __ldcg(qc + MDQ_ANALYSED);
    if (t0 >= analysed) { // this queue is handed out: help with the next one
      home = (home + 1) % mq.n;
      ++visited;
      continue;
    }
    const u32 t = t0 + lane;
    if (t >= analysed)
      continue;
    u32 c;
    const u32* slot = queue + (size_t)home * mq.cap + t;
    while ((c = ld_acquire_u32(slot)) == MD_QUEUE_EMPTY)
      __nanosleep(100);
    const u32 b = cell_start[c], e = cell_start[c + 1];
    const u32* nb = nbr + (size_t)c * MD_NBR_SLOTS;
    const u32 hdr = nb[0];
    const u32 nn = hdr & 0xffu, n_late = hdr >> 8;
    const u32 off = b; // the cell's own range of acc_xyz (one slot per point of the level): no allocation atomic
    double* own = acc_xyz + 3 * (size_t)off;
    u32 n_own = 0;
    for (u32 i = b; i < e; ++i) {
      if (state[i] != MD_UNDECIDED)
        continue;
      const double px = P[3 * (size_t)i], py = P[3 * (size_t)i + 1], pz = P[3 * (size_t)i + 2];
      bool cand = true;
      for (u32 k = 0; k < n_own && cand; ++k)
        cand = !md_in_range(px, py, pz, __ldcg(own + 3 * k), __ldcg(own + 3 * k + 1), __ldcg(own + 3 * k + 2), threshold);
      for (u32 j = 0; j < nn && cand; ++j) {
        const u64 d = __ldcg(desc + nb[1 + j]);
        const u32 cnt = (u32)(d >> 32) & 0xffffu;
        const double* src = acc_xyz + 3 * (size_t)(u32)d;
        for (u32 k = 0; k < cnt && cand; ++k)
          cand = !md_in_range(px, py, pz, __ldcg(src + 3 * k), __ldcg(src + 3 * k + 1), __ldcg(src + 3 * k + 2), threshold);
      }
      if (cand) {
        if (n_own < 0xffffu) { // the list lives in the cell's own range of acc_xyz: only the descriptor's 16 bits bound it
          __stcg(own + 3 * n_own, px);
          __stcg(own + 3 * n_own + 1, py);
          __stcg(own + 3 * n_own + 2, pz);
          ++n_own;
        } else {
          atomicExch(counters + 2, 1u); // cannot happen for cells of side < 2 spacings; reported, never silent
        }
      }
      state[i] = cand ? MD_ACCEPTED : MD_REJECTED;
    }
    __threadfence();
    __stcg(desc + c, MD_DESC_DONE | ((u64)n_own << 32) | off);
    __threadfence();
    for (u32 l = 0; l < n_late; ++l)
      md_release_later(nb[1 + nn + l], deps, queue, counters, mq);
  }
}

// one CTA of T threads per cell (dense levels: thousands of points per cell, every cell on a long chain)
#define MDW_CTA_NBR_CAP (26 * MD_OWN_CAP)
#define MDW_CTA_SMEM(T) ((MD_OWN_CAP + MDW_CTA_NBR_CAP) * 24 + (T) * 24 + (T) * 2 + (T))

template<int T>
__global__ void __launch_bounds__(T)
md_wave_cta_kernel(const double* __restrict__ P, const u32* __restrict__ cell_start, const u32* __restrict__ nbr,
                   unsigned char* __restrict__ state, u64* desc, double* acc_xyz, u32* deps, u32* queue,
                   u32* counters, double threshold, MdQueues mq)
{
  constexpr int WARPS = T / 32;
  extern __shared__ __align__(16) unsigned char md_smem[];
  double* own = reinterpret_cast<double*>(md_smem);                        // MD_OWN_CAP x 3
  double* nacc = own + MD_OWN_CAP * 3;                                     // MDW_CTA_NBR_CAP x 3
  double* s_p = nacc + MDW_CTA_NBR_CAP * 3;                                // T x 3: the batch's positions
  unsigned short* s_cand = reinterpret_cast<unsigned short*>(s_p + T * 3); // candidate threads, in order
  unsigned char* s_flag = reinterpret_cast<unsigned char*>(s_cand + T);    // accepted flag per thread
  __shared__ u32 s_c, s_total, s_ncand, s_nown, s_off;
  __shared__ u32 s_noff[32], s_ncnt[32], s_nfirst[32], s_wcount[WARPS];
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  while (true) {
    if (tid == 0)
      s_c = md_pop_cell(queue, counters, 0u, mq);
    __syncthreads();
    const u32 c = s_c;
    if (c == MD_QUEUE_EMPTY)
      break;
    const u32 b = cell_start[c], e = cell_start[c + 1];
    const u32* nb = nbr + (size_t)c * MD_NBR_SLOTS;
    const u32 nn = nb[0] & 0xffu, n_late = nb[0] >> 8;
    if (warp == 0) { // every earlier neighbour is finished: where are their accepted points
      u64 d = 0;
      if (lane < nn)
        d = __ldcg(desc + nb[1 + lane]);
      const u32 ncnt = (u32)(d >> 32) & 0xffffu;
      u32 incl = ncnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o)
          incl += t;
      }
      s_noff[lane] = (u32)d;
      s_ncnt[lane] = ncnt;
      s_nfirst[lane] = incl - ncnt;
      if (lane == 31)
        s_total = incl;
    }
    __syncthreads();
    const u32 n_nbr = s_total; // <= 26 * MD_OWN_CAP by construction
    for (u32 j = warp; j < nn; j += WARPS) // one warp copies one neighbour's accepted points
      for (u32 k = lane; k < s_ncnt[j]; k += 32) {
        const double* src = acc_xyz + 3 * (size_t)(s_noff[j] + k);
        double* dst = nacc + 3 * (s_nfirst[j] + k);
        dst[0] = __ldcg(src);
        dst[1] = __ldcg(src + 1);
        dst[2] = __ldcg(src + 2);
      }
    if (tid == 0)
      s_nown = 0;
    __syncthreads();
    for (u32 base = b; base < e; base += T) {
      const u32 i = base + tid;
      const bool valid = i < e && state[i] == MD_UNDECIDED;
      double px = 0, py = 0, pz = 0;
      if (valid) {
        px = P[3 * (size_t)i];
        py = P[3 * (size_t)i + 1];
        pz = P[3 * (size_t)i + 2];
      }
      const u32 n_checked = s_nown;
      bool cand = md_filter(valid, px, py, pz, own, n_checked, threshold);
      cand = md_filter(cand, px, py, pz, nacc, n_nbr, threshold);
      // ordered list of the surviving candidates
      const u32 wm = __ballot_sync(0xffffffffu, cand);
      if (lane == 0)
        s_wcount[warp] = __popc(wm);
      s_p[3 * tid] = px;
      s_p[3 * tid + 1] = py;
      s_p[3 * tid + 2] = pz;
      s_flag[tid] = 0;
      __syncthreads();
      u32 wbase = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w)
        wbase += (w < (int)warp) ? s_wcount[w] : 0u;
      if (cand)
        s_cand[wbase + __popc(wm & lanemask_lt())] = (unsigned short)tid;
      if (tid == T - 1)
        s_ncand = wbase + __popc(wm);
      __syncthreads();
      if (warp == 0) { // resolve in order, 32 candidates at a time
        const u32 ncand = s_ncand;
        u32 n_own = n_checked;
        for (u32 chunk = 0; chunk < ncand; chunk += 32) {
          const bool have = chunk + lane < ncand;
          const u32 t = have ? s_cand[chunk + lane] : 0u;
          const double qx = s_p[3 * t], qy = s_p[3 * t + 1], qz = s_p[3 * t + 2];
          bool cc = have;
          for (u32 k = n_checked; k < n_own; ++k) // accepted earlier in this batch
            if (cc && md_in_range(qx, qy, qz, own[3 * k], own[3 * k + 1], own[3 * k + 2], threshold))
              cc = false;
          if (md_resolve_warp(cc, qx, qy, qz, threshold, own, n_own, counters + 2))
            s_flag[t] = 1;
          __syncwarp();
        }
        if (lane == 0)
          s_nown = n_own;
      }
      __syncthreads();
      if (valid)
        state[i] = s_flag[tid] ? MD_ACCEPTED : MD_REJECTED;
    }
    const u32 n_own = s_nown;
    if (n_own) { // the cell's own range of acc_xyz (one slot per point of the level)
      if (tid == 0)
        s_off = b;
      __syncthreads();
      if (tid < n_own) {
        double* dst = acc_xyz + 3 * (size_t)(s_off + tid);
        __stcg(dst, own[3 * tid]);
        __stcg(dst + 1, own[3 * tid + 1]);
        __stcg(dst + 2, own[3 * tid + 2]);
        __threadfence();
      }
    }
    if (tid == 0)
      __stcg(desc + c, MD_DESC_DONE | ((u64)n_own << 32) | (n_own ? s_off : 0u));
    __threadfence();
    __syncthreads();
    if (tid < n_late)
      md_release_later(nb[1 + nn + tid], deps, queue, counters, mq);
  }
}

cudaError_t
run_min_distance(const SwMinDistArgs& a, SwMinDistScratch& sc, cudaStream_t stream, u32* rounds, u32* launches,
                 u64* bytes)
{
  const u64 n = a.count;
  cudaError_t e;
  const size_t tiles = sweep_tiles(n);
#define MD_TRY(x)                                                                                                      \
  do {                                                                                                                 \
    e = (x);                                                                                                           \
    if (e != cudaSuccess)                                                                                              \
      return e;                                                                                                        \
  } while (0)
  MD_TRY(grow(sc.cell_start, (n + 1) * 4));
  MD_TRY(grow(sc.cell_tile_rank0, tiles * 4));
  MD_TRY(grow(sc.state, n));
  MD_TRY(grow(sc.acc_xyz, n * 24));
  MD_TRY(grow(sc.counters, (MDQ_BASE + MD_MAX_QUEUES * MDQ_STRIDE) * 4));
  if (a.in_idx)
    MD_TRY(grow(sc.lpos, n * 24));

  // cells never span nodes: use the finer of the two prefixes
  int cell_levels = a.cell_levels;
  if (cell_levels < a.node_levels)
    cell_levels = a.node_levels;
  const int cell_shift = shift_for_levels(cell_levels);

  u32* counters = static_cast<u32*>(sc.counters.p);
  MD_TRY(cudaMemsetAsync(counters, 0, (MDQ_BASE + MD_MAX_QUEUES * MDQ_STRIDE) * 4, stream));
  launch_node_rle(a.in_key, n, cell_shift, static_cast<u32*>(sc.cell_start.p), static_cast<u32*>(sc.cell_tile_rank0.p),
                  counters + 4, sc.status, sc.ticket, stream);
  md_setup_kernel<<<(u32)tiles, SWP_THREADS, 0, stream>>>(a, static_cast<unsigned char*>(sc.state.p),
                                                         a.in_idx ? static_cast<double*>(sc.lpos.p) : nullptr,
                                                         counters + 3);
  MD_TRY(cudaMemcpyAsync(sc.h_pinned, counters + 3, 8, cudaMemcpyDeviceToHost, stream));
  MD_TRY(cudaStreamSynchronize(stream));
  const u32 n_active = sc.h_pinned[0];
  const u32 n_cells = sc.h_pinned[1];
  // (Coarser cells on the sparse levels - any cell level with side >= spacing gives the same selection - were tried
  // to cut the per-cell work of the graph: one level coarser costs 240 -> 260 ms at 125 M terrain points, two levels
  // 820 ms: a cell's points are decided one after the other, so larger cells lengthen every chain.)
  *rounds = 0;
  *launches = 2;
  *bytes = n * (8 + 8 + 1 + (a.in_idx ? 52 : 0));
  if (n_active == 0) // every node of the level is stored whole
    return cudaGetLastError();

  // table over 4 x 4 x 4 blocks of cells: a block holds at least one cell, typically 10 - 20
  u32 cap = 64;
  while ((u64)cap < (u64)n_cells + n_cells / 2 + 64) // at most 2/3 full even if every block held one cell only
    cap <<= 1;
  MD_TRY(grow(sc.hkeys, (size_t)cap * 8));
  MD_TRY(grow(sc.hvals, (size_t)cap * 4));
  MD_TRY(grow(sc.hmask, (size_t)cap * 8));
  MD_TRY(grow(sc.nbr, (size_t)n_cells * MD_NBR_SLOTS * 4));
  MD_TRY(grow(sc.desc, (size_t)n_cells * 8));
  MD_TRY(grow(sc.deps, (size_t)n_cells * 4));
  // group size by the average number of analysed points per cell: the time of one cell is on the critical path
  // of the dependency graph, so dense levels get a whole CTA per cell, the sparse ones a warp or a thread
  int mode = 0;
  static int thread_below = -1; // average analysed points per cell below which a thread takes a cell
  if (thread_below < 0) {
    const char* env = getenv("SWGPU_MD_THREAD_BELOW");
    thread_below = env ? atoi(env) : 6;
  }
  if (const char* env = getenv("SWGPU_MD_GROUP")) // tuning experiments: 1, 32, 256 or 1024
    mode = atoi(env);
  if (mode != 1 && mode != 32 && mode != 256 && mode != 1024)
    mode = (u64)n_active >= 384ull * n_cells
             ? 1024
             : ((u64)n_active >= 48ull * n_cells ? 256 : ((u64)n_active < (u64)thread_below * n_cells ? 1 : 32));
  // ready queues: one for the CTA kernels (few, heavy cells), up to MD_MAX_QUEUES for the warp / thread kernels
  MdQueues mq;
  mq.n = 1;
  if (mode == 1 || mode == 32) {
    static int max_queues = -1;
    if (max_queues < 0) {
      const char* env = getenv("SWGPU_MD_QUEUES");
      max_queues = env ? atoi(env) : MD_MAX_QUEUES;
      if (max_queues < 1 || max_queues > MD_MAX_QUEUES)
        max_queues = MD_MAX_QUEUES;
    }
    mq.n = n_cells / 16384u;
    mq.n = mq.n < 1u ? 1u : (mq.n > (u32)max_queues ? (u32)max_queues : mq.n);
  }
  mq.cap = ((n_cells + MD_QUEUE_CHUNK * mq.n - 1) / (MD_QUEUE_CHUNK * mq.n)) * MD_QUEUE_CHUNK;
  const size_t queue_entries = (size_t)mq.n * mq.cap;
  MD_TRY(grow(sc.queue, queue_entries * 4));
  MD_TRY(grow(sc.cell_active, (size_t)n_cells));
  MD_TRY(cudaMemsetAsync(sc.hkeys.p, 0, (size_t)cap * 8, stream));
  MD_TRY(cudaMemsetAsync(sc.queue.p, 0xff, queue_entries * 4, stream));
  const double* P = a.in_idx ? static_cast<const double*>(sc.lpos.p) : a.pos_sorted;
  const u32* cell_start = static_cast<const u32*>(sc.cell_start.p);
  u32* nbr = static_cast<u32*>(sc.nbr.p);
  unsigned char* state = static_cast<unsigned char*>(sc.state.p);
  unsigned char* cell_active = static_cast<unsigned char*>(sc.cell_active.p);
  u64* desc = static_cast<u64*>(sc.desc.p);
  u32* deps = static_cast<u32*>(sc.deps.p);
  u32* queue = static_cast<u32*>(sc.queue.p);
  double* acc_xyz = static_cast<double*>(sc.acc_xyz.p);
  const u32 cgrid = (n_cells + 255) / 256;
  md_hash_insert_kernel<<<cgrid, 256, 0, stream>>>(a.in_key, cell_start, n_cells, cell_shift, state, cell_active,
                                                   static_cast<u64*>(sc.hkeys.p), static_cast<u32*>(sc.hvals.p),
                                                   static_cast<u64*>(sc.hmask.p), cap - 1);
  md_neighbors_kernel<<<cgrid, 256, 0, stream>>>(a.in_key, cell_start, n_cells, cell_shift, cell_levels, a.node_levels,
                                                 cell_active, static_cast<const u64*>(sc.hkeys.p),
                                                 static_cast<const u32*>(sc.hvals.p),
                                                 static_cast<const u64*>(sc.hmask.p), cap - 1, nbr, deps, queue,
                                                 counters, mq);

  // persistent dataflow kernel: every resident group pops ready cells until all analysed cells are done
  // co-resident group counts and the dynamic shared-memory opt-in are per device
  static int warp_blocks_of[64] = {}, cta256_blocks_of[64] = {}, cta1024_blocks_of[64] = {}, thread_blocks_of[64] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  const int slot = (cur_dev >= 0 && cur_dev < 64) ? cur_dev : 0;
  int& warp_blocks = warp_blocks_of[slot];
  int& cta256_blocks = cta256_blocks_of[slot];
  int& cta1024_blocks = cta1024_blocks_of[slot];
  int& thread_blocks = thread_blocks_of[slot];
  if (!warp_blocks) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, md_wave_warp_kernel, 256, 0);
    warp_blocks = sms * (per_sm > 0 ? per_sm : 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, md_wave_thread_kernel, 256, 0);
    thread_blocks = sms * (per_sm > 0 ? per_sm : 1);
    cudaFuncSetAttribute(md_wave_cta_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, MDW_CTA_SMEM(256));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, md_wave_cta_kernel<256>, 256, MDW_CTA_SMEM(256));
    cta256_blocks = sms * (per_sm > 0 ? per_sm : 1);
    cudaFuncSetAttribute(md_wave_cta_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, MDW_CTA_SMEM(1024));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, md_wave_cta_kernel<1024>, 1024, MDW_CTA_SMEM(1024));
    cta1024_blocks = sms * (per_sm > 0 ? per_sm : 1);
  }
  if (mode == 1024) {
    const u32 blocks = n_cells < (u32)cta1024_blocks ? n_cells : (u32)cta1024_blocks;
    md_wave_cta_kernel<1024><<<blocks, 1024, MDW_CTA_SMEM(1024), stream>>>(P, cell_start, nbr, state, desc, acc_xyz,
                                                                          deps, queue, counters, a.threshold, mq);
  } else if (mode == 256) {
    const u32 blocks = n_cells < (u32)cta256_blocks ? n_cells : (u32)cta256_blocks;
    md_wave_cta_kernel<256><<<blocks, 256, MDW_CTA_SMEM(256), stream>>>(P, cell_start, nbr, state, desc, acc_xyz, deps,
                                                                        queue, counters, a.threshold, mq);
  } else if (mode == 1) {
    const u32 want = (n_cells + 255) / 256;
    const u32 blocks = want < (u32)thread_blocks ? want : (u32)thread_blocks;
    md_wave_thread_kernel<<<blocks, 256, 0, stream>>>(P, cell_start, nbr, state, desc, acc_xyz, deps, queue, counters,
                                                      a.threshold, mq);
  } else {
    const u32 want = (n_cells + 7) / 8;
    const u32 blocks = want < (u32)warp_blocks ? want : (u32)warp_blocks;
    md_wave_warp_kernel<<<blocks, 256, 0, stream>>>(P, cell_start, nbr, state, desc, acc_xyz, deps, queue, counters,
                                                    a.threshold, mq);
  }
  MD_TRY(cudaMemcpyAsync(sc.h_pinned, counters + 2, 4, cudaMemcpyDeviceToHost, stream));
  MD_TRY(cudaStreamSynchronize(stream));
  if (sc.h_pinned[0]) // a cell produced more accepted points than MD_OWN_CAP: reported, never silent
    return cudaErrorAssert;
  *rounds = 1; // swgpu_stats.min_distance_rounds now counts wavefront launches (one per sampled level)
  *launches = 5;
  *bytes = n * (8 + 8 + 1 + (a.in_idx ? 52 : 0) + 24 + 2) + (u64)n_cells * (MD_NBR_SLOTS * 4 + 8 + 28);
  return cudaGetLastError();
#undef MD_TRY
}
