// kernels_store.cu — device side of the multi-batch node store.
//
// The reference keeps what every node stores in its persistence.  When a later batch reaches a node again,
// tile_node (tiling/TilingAlgorithms.cpp:351-492) reads the node's points back (read_pnts_from_disk, :50-109:
// the node's own Morton key stays, the levels below it are re-derived with calculate_morton_index relative to
// the NODE's bounds), merges them with the incoming points (merge_node_data_sorted, tiling/Node.cpp:3-20:
// incoming first on equal keys; terminal nodes just concatenate, :22-34) and samples the union with
// AlwaysAdhereToMinSpacing (:272-275).  Points that lose their place move down with the rest.
//
// Here the store lives in HBM: all positions of all batches (global point id = position in that array) and, per
// octree level, a node table sorted by node index with the stored global ids in stored order.  One sweep level
// of a later batch then is
//   store_lookup   visited nodes (runs of the incoming list) -> slot in the level's table, stored count
//   store_fetch    the stored points of the visited nodes as a second sorted list, re-keyed like :50-109
//   merge_lists    merge path of the two lists, incoming first on ties (concat_lists for terminal levels)
//   ... the unchanged sampling kernels over the merged list (nodes with stored points never take all) ...
//   store_update   the level's table and id pool rebuilt: visited nodes take the new selection
#include "swgpu_internal.cuh"

// =============================================================================================
// exclusive scan u32 -> u64 (three kernels; the arrays are node tables, 1 .. a few million entries)
// =============================================================================================
#define SCN_THREADS 256
#define SCN_ITEMS 16
#define SCN_TILE (SCN_THREADS * SCN_ITEMS)

__device__ __forceinline__ u64
block_exclusive_u64(u64 v, u64* s_w, u64& total)
{
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u64 up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32)o)
      incl += up;
  }
  __syncthreads(); // s_w may still be read by the previous call
  if (lane == 31)
    s_w[warp] = incl;
  __syncthreads();
  u64 wofs = 0, tot = 0;
  for (u32 w = 0; w < blockDim.x / 32; ++w) {
    const u64 x = s_w[w];
    wofs += (w < warp) ? x : 0ull;
    tot += x;
  }
  total = tot;
  return wofs + incl - v;
}

__global__ void __launch_bounds__(SCN_THREADS)
scan_tile_sums_kernel(const u32* __restrict__ in, u64 n, u64* __restrict__ tile_sums)
{
  __shared__ u64 s_w[SCN_THREADS / 32];
  const u64 base = (u64)blockIdx.x * SCN_TILE + (u64)threadIdx.x * SCN_ITEMS;
  u64 sum = 0;
#pragma unroll
  for (int j = 0; j < SCN_ITEMS; ++j)
    sum += (base + j < n) ? in[base + j] : 0u;
  u64 total;
  block_exclusive_u64(sum, s_w, total);
  if (threadIdx.x == 0)
    tile_sums[blockIdx.x] = total;
}

// one block: tile_sums -> exclusive offsets in place, grand total -> *total_out
__global__ void __launch_bounds__(1024)
scan_tile_offsets_kernel(u64* __restrict__ tile_sums, u32 n_tiles, u64* __restrict__ total_out)
{
  __shared__ u64 s_w[32];
  __shared__ u64 s_carry;
  if (threadIdx.x == 0)
    s_carry = 0;
  __syncthreads();
  for (u32 t0 = 0; t0 < n_tiles; t0 += 1024) {
    const u32 t = t0 + threadIdx.x;
    const u64 v = t < n_tiles ? tile_sums[t] : 0ull;
    u64 total;
    const u64 excl = block_exclusive_u64(v, s_w, total);
    const u64 carry = s_carry;
    if (t < n_tiles)
      tile_sums[t] = carry + excl;
    __syncthreads();
    if (threadIdx.x == 0)
      s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *total_out = s_carry;
}

__global__ void __launch_bounds__(SCN_THREADS)
scan_apply_kernel(const u32* __restrict__ in, u64 n, const u64* __restrict__ tile_offs, const u64* __restrict__ total,
                  u64* __restrict__ out)
{
  __shared__ u64 s_w[SCN_THREADS / 32];
  const u64 base = (u64)blockIdx.x * SCN_TILE + (u64)threadIdx.x * SCN_ITEMS;
  u32 v[SCN_ITEMS];
  u64 sum = 0;
#pragma unroll
  for (int j = 0; j < SCN_ITEMS; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0u;
    sum += v[j];
  }
  u64 tot;
  u64 run = tile_offs[blockIdx.x] + block_exclusive_u64(sum, s_w, tot);
#pragma unroll
  for (int j = 0; j < SCN_ITEMS; ++j) {
    if (base + j < n)
      out[base + j] = run;
    run += v[j];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    out[n] = *total;
}

size_t
scan_scratch_words(u64 n)
{
  return (size_t)((n + SCN_TILE - 1) / SCN_TILE) + 2;
}

// out[0..n] = exclusive scan of in[0..n) (out[n] = total).  scratch: scan_scratch_words(n) u64.
void
launch_exclusive_scan_u32(const u32* in, u64 n, u64* out, u64* scratch, cudaStream_t stream)
{
  const u32 tiles = (u32)((n + SCN_TILE - 1) / SCN_TILE);
  u64* total = scratch + tiles;
  if (tiles == 0) {
    cudaMemsetAsync(out, 0, sizeof(u64), stream);
    return;
  }
  scan_tile_sums_kernel<<<tiles, SCN_THREADS, 0, stream>>>(in, n, scratch);
  scan_tile_offsets_kernel<<<1, 1024, 0, stream>>>(scratch, tiles, total);
  scan_apply_kernel<<<tiles, SCN_THREADS, 0, stream>>>(in, n, scratch, total, out);
}

// =============================================================================================
// small helpers
// =============================================================================================
__device__ __forceinline__ u32
lower_bound_u64(const u64* __restrict__ a, u32 n, u64 v)
{
  u32 lo = 0, hi = n;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if (a[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

// last k in [0, n) with a[k] <= v (a[0] <= v is guaranteed by the callers)
__device__ __forceinline__ u32
owner_u64(const u64* __restrict__ a, u32 n, u64 v)
{
  u32 lo = 0, hi = n;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if (a[mid] <= v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo - 1;
}

__device__ __forceinline__ u32
owner_u32(const u32* __restrict__ a, u32 n, u32 v)
{
  u32 lo = 0, hi = n;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if (a[mid] <= v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo - 1;
}

// gid[i] = base + order[i]: sorted position -> global point id of the batch that starts at `base`
__global__ void __launch_bounds__(256)
make_gids_kernel(const u32* __restrict__ order, u64 n, u32 base, u32* __restrict__ gid)
{
  const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
  if (i < n)
    gid[i] = base + order[i];
}

void
launch_make_gids(const u32* order, u64 n, u32 base, u32* gid, cudaStream_t stream)
{
  if (n)
    make_gids_kernel<<<(u32)((n + 255) / 256), 256, 0, stream>>>(order, n, base, gid);
}

// =============================================================================================
// visited nodes -> slots of the level's table
// =============================================================================================
__global__ void __launch_bounds__(256)
store_lookup_kernel(const u64* __restrict__ in_key, const u32* __restrict__ node_start, u32 n_nodes, int node_shift,
                    const u64* __restrict__ st_index, u32 st_n, const u64* __restrict__ st_first,
                    u32* __restrict__ slot, u32* __restrict__ cnt)
{
  const u32 k = blockIdx.x * 256 + threadIdx.x;
  if (k >= n_nodes)
    return;
  const u64 idx = (in_key[node_start[k]] & SW_KEY_MASK) >> node_shift;
  const u32 lb = lower_bound_u64(st_index, st_n, idx);
  const bool found = lb < st_n && st_index[lb] == idx;
  slot[k] = found ? lb : 0xFFFFFFFFu;
  cnt[k] = found ? (u32)(st_first[lb + 1] - st_first[lb]) : 0u;
}

void
launch_store_lookup(const u64* in_key, const u32* node_start, u32 n_nodes, int node_shift, const u64* st_index,
                    u32 st_n, const u64* st_first, u32* slot, u32* cnt, cudaStream_t stream)
{
  if (n_nodes)
    store_lookup_kernel<<<(n_nodes + 255) / 256, 256, 0, stream>>>(in_key, node_start, n_nodes, node_shift, st_index,
                                                                  st_n, st_first, slot, cnt);
}

// =============================================================================================
// stored points of the visited nodes as a sorted (key, gid) list
// =============================================================================================
// read_pnts_from_disk, TilingAlgorithms.cpp:50-109: `key = node key with the levels below the node taken from
// calculate_morton_index(position, NODE bounds)`; the node bounds come from the get_octant_bounds recurrence
// (get_bounds_from_node_index), the scale is 2^21 / extent of THAT box (OctreeAlgorithms.h:69).
__global__ void __launch_bounds__(256)
store_fetch_kernel(u64 m, const u64* __restrict__ boff, u32 n_nodes, const u32* __restrict__ slot,
                   const u64* __restrict__ st_first, const u32* __restrict__ st_ids, const u64* __restrict__ in_key,
                   const u32* __restrict__ node_start, int levels, const double* __restrict__ xyz, SwBounds root,
                   u64* __restrict__ bkey, u32* __restrict__ bidx)
{
  const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
  if (j >= m)
    return;
  const u32 k = owner_u64(boff, n_nodes, j);
  const u32 gid = st_ids[st_first[slot[k]] + (j - boff[k])];
  const u64 node_key = in_key[node_start[k]] & SW_KEY_MASK;
  SwBounds nb;
  bounds_from_key(node_key, levels, root, nb.min, nb.max);
#pragma unroll
  for (int a = 0; a < 3; ++a)
    nb.scale[a] = 2097152.0 / (nb.max[a] - nb.min[a]);
  const double* p = xyz + 3 * (u64)gid;
  const u64 below = morton_from_position(p[0], p[1], p[2], nb);
  const int node_shift = shift_for_levels(levels);
  const u64 prefix = levels ? ((node_key >> node_shift) << node_shift) : 0ull;
  bkey[j] = prefix | (levels < 21 ? (below >> (3 * levels)) : 0ull);
  bidx[j] = gid;
}

void
launch_store_fetch(u64 m, const u64* boff, u32 n_nodes, const u32* slot, const u64* st_first, const u32* st_ids,
                   const u64* in_key, const u32* node_start, int levels, const double* xyz, const SwBounds& root,
                   u64* bkey, u32* bidx, cudaStream_t stream)
{
  if (m)
    store_fetch_kernel<<<(u32)((m + 255) / 256), 256, 0, stream>>>(m, boff, n_nodes, slot, st_first, st_ids, in_key,
                                                                  node_start, levels, xyz, root, bkey, bidx);
}

// keys of stored ids relative to the ROOT bounds (reconstruct: index_point over the children's points,
// TilingAlgorithms.cpp:1661-1691; the stored positions are clamped already)
__global__ void __launch_bounds__(256)
store_root_keys_kernel(const u32* __restrict__ ids, u64 n, const double* __restrict__ xyz, SwBounds root,
                       u64* __restrict__ keys)
{
  const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
  if (i >= n)
    return;
  const double* p = xyz + 3 * (u64)ids[i];
  keys[i] = morton_from_position(p[0], p[1], p[2], root);
}

void
launch_store_root_keys(const u32* ids, u64 n, const double* xyz, const SwBounds& root, u64* keys, cudaStream_t stream)
{
  if (n)
    store_root_keys_kernel<<<(u32)((n + 255) / 256), 256, 0, stream>>>(ids, n, xyz, root, keys);
}

// node boundaries of the merged list + the per-node count the take-all test sees: a node that has stored
// points is always sampled (AlwaysAdhereToMinSpacing, TilingAlgorithms.cpp:272-275) -> count = 2^32 - 1
__global__ void __launch_bounds__(256)
store_merged_nodes_kernel(const u32* __restrict__ node_start_a, const u64* __restrict__ boff, u32 n_nodes,
                          u32* __restrict__ node_start_c, u32* __restrict__ gcount)
{
  const u32 k = blockIdx.x * 256 + threadIdx.x;
  if (k > n_nodes)
    return;
  node_start_c[k] = node_start_a[k] + (u32)boff[k];
  if (k < n_nodes)
    gcount[k] = (boff[k + 1] > boff[k]) ? 0xFFFFFFFFu : node_start_a[k + 1] - node_start_a[k];
}

void
launch_store_merged_nodes(const u32* node_start_a, const u64* boff, u32 n_nodes, u32* node_start_c, u32* gcount,
                          cudaStream_t stream)
{
  store_merged_nodes_kernel<<<(n_nodes + 1 + 255) / 256, 256, 0, stream>>>(node_start_a, boff, n_nodes, node_start_c,
                                                                           gcount);
}

// =============================================================================================
// merge path: C = merge(A, B), A first on equal keys (std::merge(incoming, cached), Node.cpp:3-20)
// =============================================================================================
#define MRG_THREADS 256
#define MRG_ITEMS 8
#define MRG_TILE (MRG_THREADS * MRG_ITEMS)

// number of A elements among the first `diag` outputs
template<typename KA, typename KB>
__device__ __forceinline__ u32
merge_path(KA a_at, u32 na, KB b_at, u32 nb, u32 diag)
{
  u32 lo = diag > nb ? diag - nb : 0u;
  u32 hi = diag < na ? diag : na;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if (a_at(mid) <= b_at(diag - 1 - mid)) // A[mid] goes before that B element
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(MRG_THREADS)
merge_lists_kernel(const u64* __restrict__ ak, const u32* __restrict__ ai, u32 na, const u64* __restrict__ bk,
                   const u32* __restrict__ bi, u32 nb, u64* __restrict__ ck, u32* __restrict__ ci)
{
  __shared__ u64 s_k[MRG_TILE];
  __shared__ u32 s_i[MRG_TILE];
  __shared__ u32 s_split[2];
  const u32 tid = threadIdx.x;
  const u64 total = (u64)na + nb;
  const u64 d0 = (u64)blockIdx.x * MRG_TILE;
  const u64 d1 = d0 + MRG_TILE < total ? d0 + MRG_TILE : total;
  if (tid < 2) {
    const u64 d = tid ? d1 : d0;
    s_split[tid] = merge_path([&](u32 i) { return ak[i] & SW_KEY_MASK; }, na, [&](u32 i) { return bk[i] & SW_KEY_MASK; },
                              nb, (u32)d);
  }
  __syncthreads();
  const u32 a0 = s_split[0], a1 = s_split[1];
  const u32 b0 = (u32)d0 - a0, b1 = (u32)d1 - a1;
  const u32 la = a1 - a0, lb = b1 - b0;
  // stage A part then B part
  for (u32 p = tid; p < la + lb; p += MRG_THREADS) {
    if (p < la) {
      s_k[p] = ak[a0 + p] & SW_KEY_MASK;
      s_i[p] = ai[a0 + p];
    } else {
      s_k[p] = bk[b0 + p - la] & SW_KEY_MASK;
      s_i[p] = bi[b0 + p - la];
    }
  }
  __syncthreads();
  const u32 n_out = la + lb;
  const u32 diag = tid * MRG_ITEMS < n_out ? tid * MRG_ITEMS : n_out;
  const u64* sa = s_k;
  const u64* sb = s_k + la;
  u32 ia = merge_path([&](u32 i) { return sa[i]; }, la, [&](u32 i) { return sb[i]; }, lb, diag);
  u32 ib = diag - ia;
  u64 ok[MRG_ITEMS];
  u32 oi[MRG_ITEMS];
#pragma unroll
  for (int j = 0; j < MRG_ITEMS; ++j) {
    const u32 o = diag + j;
    if (o < n_out) {
      const bool take_a = ia < la && (ib >= lb || sa[ia] <= sb[ib]);
      const u32 src = take_a ? ia : la + ib;
      ok[j] = s_k[src];
      oi[j] = s_i[src];
      ia += take_a ? 1u : 0u;
      ib += take_a ? 0u : 1u;
    }
  }
  __syncthreads(); // every thread has read its inputs: the staging arrays become the output tile
#pragma unroll
  for (int j = 0; j < MRG_ITEMS; ++j) {
    const u32 o = diag + j;
    if (o < n_out) {
      s_k[o] = ok[j];
      s_i[o] = oi[j];
    }
  }
  __syncthreads();
  for (u32 p = tid; p < n_out; p += MRG_THREADS) {
    ck[d0 + p] = s_k[p];
    ci[d0 + p] = s_i[p];
  }
}

void
launch_merge_lists(const u64* ak, const u32* ai, u64 na, const u64* bk, const u32* bi, u64 nb, u64* ck, u32* ci,
                   cudaStream_t stream)
{
  const u64 total = na + nb;
  if (total)
    merge_lists_kernel<<<(u32)((total + MRG_TILE - 1) / MRG_TILE), MRG_THREADS, 0, stream>>>(ak, ai, (u32)na, bk, bi,
                                                                                           (u32)nb, ck, ci);
}

// terminal nodes: merge_node_data_unsorted (Node.cpp:22-34) = incoming points, then the stored ones
__global__ void __launch_bounds__(256)
concat_lists_kernel(const u64* __restrict__ ak, const u32* __restrict__ ai, u32 na, const u64* __restrict__ bk,
                    const u32* __restrict__ bi, u32 nb, const u32* __restrict__ node_start_a,
                    const u64* __restrict__ boff, u32 n_nodes, u64* __restrict__ ck, u32* __restrict__ ci)
{
  const u64 t = (u64)blockIdx.x * 256 + threadIdx.x;
  if (t < na) {
    const u32 k = owner_u32(node_start_a, n_nodes, (u32)t);
    const u64 dst = t + boff[k];
    ck[dst] = ak[t] & SW_KEY_MASK;
    ci[dst] = ai[t];
  } else if (t < (u64)na + nb) {
    const u64 j = t - na;
    const u32 k = owner_u64(boff, n_nodes, j);
    const u64 dst = (u64)node_start_a[k + 1] + j;
    ck[dst] = bk[j];
    ci[dst] = bi[j];
  }
}

void
launch_concat_lists(const u64* ak, const u32* ai, u64 na, const u64* bk, const u32* bi, u64 nb,
                    const u32* node_start_a, const u64* boff, u32 n_nodes, u64* ck, u32* ci, cudaStream_t stream)
{
  const u64 total = na + nb;
  if (total)
    concat_lists_kernel<<<(u32)((total + 255) / 256), 256, 0, stream>>>(ak, ai, (u32)na, bk, bi, (u32)nb, node_start_a,
                                                                      boff, n_nodes, ck, ci);
}

// =============================================================================================
// store update: the level's table = (old nodes that were not visited) U (visited nodes), sorted by index
// =============================================================================================
// visited node j -> position of its index in the old table + whether it exists there
__global__ void __launch_bounds__(256)
store_match_kernel(const u64* __restrict__ vidx, u32 nv, const u64* __restrict__ oidx, u32 no, u32* __restrict__ lo,
                   u32* __restrict__ found)
{
  const u32 j = blockIdx.x * 256 + threadIdx.x;
  if (j >= nv)
    return;
  const u32 lb = lower_bound_u64(oidx, no, vidx[j]);
  lo[j] = lb;
  found[j] = (lb < no && oidx[lb] == vidx[j]) ? 1u : 0u;
}

// rows of the new table: index, count, flags, source (bit 63: the batch's output chunk, else the old pool)
__global__ void __launch_bounds__(256)
store_rows_kernel(const u64* __restrict__ vidx, const u64* __restrict__ vfirst, u32 nv, u64 chunk_end, u32 chunk_flags,
                  const u32* __restrict__ lo, const u64* __restrict__ cumf, const u64* __restrict__ oidx,
                  const u64* __restrict__ ofirst, const u32* __restrict__ oflags, u32 no, u64* __restrict__ nidx,
                  u32* __restrict__ ncnt, u32* __restrict__ nflags, u64* __restrict__ nsrc)
{
  const u32 t = blockIdx.x * 256 + threadIdx.x;
  if (t < nv) {
    const u32 pos = t + lo[t] - (u32)cumf[t];
    const u64 fmask = ~(1ull << 63);
    const u64 f = vfirst[t] & fmask;
    const u64 e = (t + 1 < nv) ? (vfirst[t + 1] & fmask) : chunk_end;
    nidx[pos] = vidx[t];
    ncnt[pos] = (u32)(e - f);
    nflags[pos] = chunk_flags | ((vfirst[t] >> 63) ? SW_NODE_TAKE_ALL : 0u);
    nsrc[pos] = (1ull << 63) | f;
  } else if (t < nv + no) {
    const u32 i = t - nv;
    const u32 lb = lower_bound_u64(vidx, nv, oidx[i]);
    if (lb < nv && vidx[lb] == oidx[i])
      return; // replaced by the visit
    const u32 pos = i + lb - (u32)cumf[lb];
    nidx[pos] = oidx[i];
    ncnt[pos] = (u32)(ofirst[i + 1] - ofirst[i]);
    nflags[pos] = oflags[i];
    nsrc[pos] = ofirst[i];
  }
}

__global__ void __launch_bounds__(256)
store_copy_kernel(u64 total, const u64* __restrict__ nfirst, u32 nn, const u64* __restrict__ nsrc,
                  const u32* __restrict__ old_ids, const u32* __restrict__ chunk_ids, u32* __restrict__ new_ids)
{
  const u64 e = (u64)blockIdx.x * 256 + threadIdx.x;
  if (e >= total)
    return;
  const u32 k = owner_u64(nfirst, nn, e);
  const u64 src = nsrc[k];
  const u64 off = e - nfirst[k];
  new_ids[e] = (src >> 63) ? chunk_ids[(src & ~(1ull << 63)) + off] : old_ids[src + off];
}

void
launch_store_match(const u64* vidx, u32 nv, const u64* oidx, u32 no, u32* lo, u32* found, cudaStream_t stream)
{
  if (nv)
    store_match_kernel<<<(nv + 255) / 256, 256, 0, stream>>>(vidx, nv, oidx, no, lo, found);
}

void
launch_store_rows(const u64* vidx, const u64* vfirst, u32 nv, u64 chunk_end, u32 chunk_flags, const u32* lo,
                  const u64* cumf, const u64* oidx, const u64* ofirst, const u32* oflags, u32 no, u64* nidx, u32* ncnt,
                  u32* nflags, u64* nsrc, cudaStream_t stream)
{
  if (nv + no)
    store_rows_kernel<<<(nv + no + 255) / 256, 256, 0, stream>>>(vidx, vfirst, nv, chunk_end, chunk_flags, lo, cumf,
                                                                oidx, ofirst, oflags, no, nidx, ncnt, nflags, nsrc);
}

void
launch_store_copy(u64 total, const u64* nfirst, u32 nn, const u64* nsrc, const u32* old_ids, const u32* chunk_ids,
                  u32* new_ids, cudaStream_t stream)
{
  if (total)
    store_copy_kernel<<<(u32)((total + 255) / 256), 256, 0, stream>>>(total, nfirst, nn, nsrc, old_ids, chunk_ids,
                                                                     new_ids);
}
