// kernels_shard.cu — device side of the multi-GPU path: every GPU ends up owning
// whole Morton-prefix subtrees.
//
// The reference is a single process (no collective exists in it); its unit of independent work is
// the octree node (do_tiling_for_node, tiling/TilingAlgorithms.cpp:499-561, one task per start node
// at :1314-1351).  Sharding by the leading `shard_levels` octree levels of the Morton key keeps
// that unit intact:
//
//   prefix_histogram      counts of the 8^6 level-5 prefixes of UNSORTED keys.  Summed over the
//                         ranks it yields (a) FAST's start level exactly as
//                         estimate_start_node_level_in_octree computes it (:1473-1535) and (b) the
//                         splitters: contiguous prefix ranges of equal point count
//   partition_by_splitter stable multi-way partition of (position, global id) by destination rank
//                         = the send buffer of the all-to-all.  Stability + rank-ordered receive
//                         keeps points in global-id order on the receiver, so the receiver's stable
//                         sort reproduces the single-GPU tie rule (key, original index)
//   node_count exchange   nodes above the shard prefix depth span GPUs; their take-all decision
//                         (Sampling.h:201-208) needs the global point count: dense per-prefix
//                         counters are filled here, summed by the caller's collective, read back
#include "swgpu_internal.cuh"

// ---------------------------------------------------------------------------------------------
// level-5 prefix histogram of unsorted keys
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prefix_histogram_kernel(const u64* __restrict__ keys, u64 n, u32* __restrict__ bins)
{
  // one L2 reduction per key (no return value); __match_any_sync pre-aggregation costs more than it
  // saves on B200 (~60 SM-cycles per warp instruction)
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256)
    atomicAdd(&bins[(u32)((keys[i] & SW_KEY_MASK) >> 45)], 1u);
}

void
launch_prefix_histogram(const u64* keys, u64 n, u32* bins, cudaStream_t stream)
{
  if (n == 0)
    return;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  u64 want = (n + 255) / 256;
  const u64 cap = (u64)sms * 8;
  prefix_histogram_kernel<<<(u32)(want < cap ? want : cap), 256, 0, stream>>>(keys, n, bins);
}

// Coarse prefix histogram (<= 4 octree levels = 4096 bins): privatised in shared memory, so the cost is one
// read of the keys instead of one L2 atomic per key (the exact 8^6-bin histogram above is atomics-bound:
// 1.8 ms per 100 M keys on B200).  Enough for balanced splitters; the exact level-5 counts FAST's start
// level needs are taken from the SORTED keys after the exchange (bin boundaries by binary search).
__global__ void __launch_bounds__(256)
prefix_histogram_coarse_kernel(const u64* __restrict__ keys, u64 n, int shift, u32 n_bins, u32* __restrict__ bins)
{
  __shared__ u32 s_bins[4096];
  for (u32 i = threadIdx.x; i < n_bins; i += 256)
    s_bins[i] = 0;
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256)
    atomicAdd(&s_bins[(u32)((keys[i] & SW_KEY_MASK) >> shift)], 1u);
  __syncthreads();
  for (u32 i = threadIdx.x; i < n_bins; i += 256) {
    const u32 v = s_bins[i];
    if (v)
      atomicAdd(&bins[i], v);
  }
}

void
launch_prefix_histogram_coarse(const u64* keys, u64 n, int levels, u32* bins, cudaStream_t stream)
{
  if (n == 0)
    return;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  u64 want = (n + 256 * 16 - 1) / (256 * 16);
  const u64 cap = (u64)sms * 8;
  prefix_histogram_coarse_kernel<<<(u32)(want < cap ? (want ? want : 1) : cap), 256, 0, stream>>>(
    keys, n, 63 - 3 * levels, 1u << (3 * levels), bins);
}

// counts[b] = bin_start[b + 1] - bin_start[b]
__global__ void __launch_bounds__(256)
bin_counts_kernel(const u32* __restrict__ bin_start, u32 n_bins, u32* __restrict__ counts)
{
  const u32 b = blockIdx.x * 256 + threadIdx.x;
  if (b < n_bins)
    counts[b] = bin_start[b + 1] - bin_start[b];
}

void
launch_bin_counts(const u32* bin_start, u32 n_bins, u32* counts, cudaStream_t stream)
{
  bin_counts_kernel<<<(n_bins + 255) / 256, 256, 0, stream>>>(bin_start, n_bins, counts);
}

// ---------------------------------------------------------------------------------------------
// stable multi-way partition by destination rank
// ---------------------------------------------------------------------------------------------
#define PT_THREADS 256
#define PT_WARPS 8
#ifndef PT_ITEMS
#define PT_ITEMS 4
#endif
#define PT_TILE (PT_THREADS * PT_ITEMS)

struct SwSplitters
{
  u32 first_prefix[SW_MAX_RANKS]; // rank r owns level-5 prefixes [first_prefix[r], first_prefix[r+1])
  u32 n_ranks;
};

__device__ __forceinline__ u32
dest_of(u64 key, const SwSplitters& sp)
{
  const u32 prefix = (u32)((key & SW_KEY_MASK) >> 45);
  u32 d = 0;
#pragma unroll
  for (u32 r = 1; r < SW_MAX_RANKS; ++r)
    d += (r < sp.n_ranks && prefix >= sp.first_prefix[r]) ? 1u : 0u;
  return d;
}

// tile_counts[tile * SW_MAX_RANKS + d] = points of the tile that go to rank d
__global__ void __launch_bounds__(PT_THREADS)
partition_count_kernel(const u64* __restrict__ keys, u64 n, SwSplitters sp, u32* __restrict__ tile_counts)
{
  __shared__ u32 s_cnt[SW_MAX_RANKS];
  if (threadIdx.x < SW_MAX_RANKS)
    s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const u64 base = (u64)blockIdx.x * PT_TILE;
  u32 local[SW_MAX_RANKS];
#pragma unroll
  for (u32 r = 0; r < SW_MAX_RANKS; ++r)
    local[r] = 0;
#pragma unroll
  for (int j = 0; j < PT_ITEMS; ++j) {
    const u64 i = base + j * PT_THREADS + threadIdx.x;
    if (i < n) {
      const u32 d = dest_of(keys[i], sp);
#pragma unroll
      for (u32 r = 0; r < SW_MAX_RANKS; ++r)
        local[r] += (d == r) ? 1u : 0u;
    }
  }
#pragma unroll
  for (u32 r = 0; r < SW_MAX_RANKS; ++r) {
    u32 v = local[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v)
      atomicAdd(&s_cnt[r], v);
  }
  __syncthreads();
  if (threadIdx.x < SW_MAX_RANKS)
    tile_counts[(u64)blockIdx.x * SW_MAX_RANKS + threadIdx.x] = s_cnt[threadIdx.x];
}

// One block per destination: exclusive scan of its column over the tiles (in place), total to
// send_counts[d].
__global__ void __launch_bounds__(1024)
partition_scan_kernel(u32* __restrict__ tile_counts, u32 n_tiles, u64* __restrict__ send_counts)
{
  __shared__ u32 s_warp[32];
  __shared__ u32 s_carry;
  const u32 d = blockIdx.x;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    s_carry = 0;
  __syncthreads();
  for (u32 t0 = 0; t0 < n_tiles; t0 += 1024) {
    const u32 t = t0 + threadIdx.x;
    const u32 v = (t < n_tiles) ? tile_counts[(u64)t * SW_MAX_RANKS + d] : 0u;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (u32)o)
        incl += up;
    }
    if (lane == 31)
      s_warp[warp] = incl;
    __syncthreads();
    u32 wofs = 0;
    for (u32 w = 0; w < warp; ++w)
      wofs += s_warp[w];
    const u32 carry = s_carry;
    if (t < n_tiles)
      tile_counts[(u64)t * SW_MAX_RANKS + d] = carry + wofs + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023)
      s_carry = carry + wofs + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0)
    send_counts[d] = s_carry;
}

// Scatter.  Position of a point = base of (this source, destination d) + offset of the tile + rank
// inside the tile (tile order = index order, so the partition is stable).  The tile is staged in
// shared memory ordered by destination, then every destination's segment leaves as one contiguous,
// coalesced write.  The destination bases are plain device pointers: slices of a local send buffer
// (NCCL all-to-all follows) or PEER memory mapped over NVLink (partition + exchange in one kernel,
// no send buffer, no collective on the data path).
struct SwPartitionDst
{
  double* xyz[SW_MAX_RANKS]; // where this source's points for destination d start
  u32* ids[SW_MAX_RANKS];
  u32* attr[SW_MAX_RANKS];   // optional per-point attribute record (AW 32-bit words), same order as the points
};

// AW: 32-bit words of the attribute record that travels with every point (0 = none; LAS point format 2 carries
// 14 attribute bytes per point, core/datastructures/PointBuffer.h:291-304, packed into 16)
template<int AW>
__global__ void __launch_bounds__(PT_THREADS)
partition_scatter_kernel(const u64* __restrict__ keys, const double* __restrict__ xyz, const u32* __restrict__ attr,
                         u64 n, SwSplitters sp, const u32* __restrict__ tile_offsets,
                         const u64* __restrict__ send_counts, u32 id_base, SwPartitionDst dst)
{
  __shared__ double s_x[3 * PT_TILE];
  __shared__ u32 s_id[PT_TILE];
  __shared__ u32 s_at[AW ? AW * PT_TILE : 1];
  __shared__ u32* s_ga[SW_MAX_RANKS];
  __shared__ u32 s_wcnt[PT_WARPS][SW_MAX_RANKS];
  __shared__ u32 s_wbase[PT_WARPS][SW_MAX_RANKS]; // first staged position of (warp, destination)
  __shared__ u32 s_tot[SW_MAX_RANKS];
  __shared__ u32 s_seg[SW_MAX_RANKS + 1]; // first staged position of every destination's segment
  __shared__ double* s_gx[SW_MAX_RANKS];  // destination pointers, biased by the segment start
  __shared__ u32* s_gi[SW_MAX_RANKS];
  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const u64 base = (u64)blockIdx.x * PT_TILE;
  const u32 lt = lanemask_lt();
  // warp-striped: item j of a warp covers 32 consecutive points; a warp owns PT_ITEMS * 32 of them.
  // One ballot per (item, destination) yields both the warp's counts and the point's rank inside its
  // (warp, destination) group: items in order, lanes in order = index order, so the partition is stable.
  u32 dest[PT_ITEMS];
  u32 low[PT_ITEMS];
  u32 cnt[SW_MAX_RANKS];
#pragma unroll
  for (u32 r = 0; r < SW_MAX_RANKS; ++r)
    cnt[r] = 0;
#pragma unroll
  for (int j = 0; j < PT_ITEMS; ++j) {
    const u64 i = base + warp * (32 * PT_ITEMS) + j * 32 + lane;
    dest[j] = (i < n) ? dest_of(keys[i], sp) : 0xFFu;
    low[j] = 0;
#pragma unroll
    for (u32 r = 0; r < SW_MAX_RANKS; ++r)
      if (r < sp.n_ranks) { // warp-uniform: a ballot costs ~2.6 SM-cycles, skip the unused ranks
        const u32 m = __ballot_sync(0xffffffffu, dest[j] == r);
        if (dest[j] == r)
          low[j] = cnt[r] + __popc(m & lt);
        cnt[r] += __popc(m);
      }
  }
  if (lane < SW_MAX_RANKS) {
    u32 mine = 0;
#pragma unroll
    for (u32 r = 0; r < SW_MAX_RANKS; ++r)
      mine = (lane == r) ? cnt[r] : mine;
    s_wcnt[warp][lane] = mine;
  }
  __syncthreads();
  if (tid < SW_MAX_RANKS) {
    u32 tot = 0;
#pragma unroll
    for (u32 w = 0; w < PT_WARPS; ++w)
      tot += s_wcnt[w][tid];
    s_tot[tid] = tot;
  }
  __syncthreads();
  if (tid < SW_MAX_RANKS) {
    const u32 r = tid;
    u32 seg = 0;
    u64 sent_before = 0; // local send buffer: destinations are laid out one after the other
    for (u32 q = 0; q < r; ++q) {
      seg += s_tot[q];
      if (send_counts && q < sp.n_ranks)
        sent_before += send_counts[q];
    }
    s_seg[r] = seg;
    if (r == SW_MAX_RANKS - 1)
      s_seg[SW_MAX_RANKS] = seg + s_tot[r];
    if (r < sp.n_ranks) {
      const u64 off = sent_before + tile_offsets[(u64)blockIdx.x * SW_MAX_RANKS + r];
      // staged element e of this destination's segment goes to pointer[e]: bias by the segment start
      s_gx[r] = dst.xyz[r] + 3 * off - 3 * (u64)seg;
      s_gi[r] = dst.ids[r] + off - seg;
      if (AW)
        s_ga[r] = dst.attr[r] + (u64)AW * off - (u64)AW * seg;
    }
  }
  __syncthreads();
  if (tid < PT_WARPS * SW_MAX_RANKS) {
    const u32 w = tid / SW_MAX_RANKS, r = tid % SW_MAX_RANKS;
    u32 o = s_seg[r];
    for (u32 q = 0; q < w; ++q)
      o += s_wcnt[q][r];
    s_wbase[w][r] = o;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PT_ITEMS; ++j) {
    const u64 i = base + warp * (32 * PT_ITEMS) + j * 32 + lane;
    if (i < n) {
      const u32 pos = s_wbase[warp][dest[j]] + low[j];
      s_x[3 * pos] = xyz[3 * i];
      s_x[3 * pos + 1] = xyz[3 * i + 1];
      s_x[3 * pos + 2] = xyz[3 * i + 2];
      s_id[pos] = id_base + (u32)i;
#pragma unroll
      for (int w = 0; w < AW; ++w)
        s_at[AW * pos + w] = attr[(u64)AW * i + w];
    }
  }
  __syncthreads();
  // copy-out, one destination after the other: a segment leaves as one contiguous run of coalesced stores
  // (no per-element destination lookup, no division)
  for (u32 r = 0; r < sp.n_ranks; ++r) {
    const u32 b = s_seg[r], e = s_seg[r + 1];
    double* gx = s_gx[r];
    u32* gi = s_gi[r];
    for (u32 k = 3 * b + tid; k < 3 * e; k += PT_THREADS)
      gx[k] = s_x[k];
    for (u32 k = b + tid; k < e; k += PT_THREADS)
      gi[k] = s_id[k];
    if (AW) {
      u32* ga = s_ga[r];
      for (u32 k = AW * b + tid; k < AW * e; k += PT_THREADS)
        ga[k] = s_at[k];
    }
  }
}

static void
launch_partition_scatter(u32 tiles, const u64* keys, const double* xyz, const u32* attr, u32 attr_words, u64 n,
                         const SwSplitters& sp, const u32* tile_counts, const u64* send_counts, u32 id_base,
                         const SwPartitionDst& dst, cudaStream_t stream)
{
  switch (attr ? attr_words : 0u) {
    case 0:
      partition_scatter_kernel<0><<<tiles, PT_THREADS, 0, stream>>>(keys, xyz, nullptr, n, sp, tile_counts, send_counts, id_base, dst);
      break;
    case 1:
      partition_scatter_kernel<1><<<tiles, PT_THREADS, 0, stream>>>(keys, xyz, attr, n, sp, tile_counts, send_counts, id_base, dst);
      break;
    case 2:
      partition_scatter_kernel<2><<<tiles, PT_THREADS, 0, stream>>>(keys, xyz, attr, n, sp, tile_counts, send_counts, id_base, dst);
      break;
    case 3:
      partition_scatter_kernel<3><<<tiles, PT_THREADS, 0, stream>>>(keys, xyz, attr, n, sp, tile_counts, send_counts, id_base, dst);
      break;
    default:
      partition_scatter_kernel<4><<<tiles, PT_THREADS, 0, stream>>>(keys, xyz, attr, n, sp, tile_counts, send_counts, id_base, dst);
      break;
  }
}

size_t
partition_tiles(u64 n)
{
  const size_t t = (size_t)((n + PT_TILE - 1) / PT_TILE);
  return t ? t : 1;
}

static SwSplitters
make_splitters(const u32* first_prefix, u32 n_ranks)
{
  SwSplitters sp{};
  sp.n_ranks = n_ranks;
  for (u32 r = 0; r < SW_MAX_RANKS; ++r)
    sp.first_prefix[r] = r < n_ranks ? first_prefix[r] : 0xFFFFFFFFu;
  return sp;
}

void
launch_partition_by_splitters(const u64* keys, const double* xyz, u64 n, const u32* first_prefix, u32 n_ranks,
                              u32 id_base, u32* tile_counts, u64* send_counts, double* out_xyz, u32* out_id,
                              const u32* attr, u32 attr_words, u32* out_attr, cudaStream_t stream)
{
  const SwSplitters sp = make_splitters(first_prefix, n_ranks);
  const u32 tiles = (u32)partition_tiles(n);
  if (n == 0) {
    cudaMemsetAsync(send_counts, 0, SW_MAX_RANKS * sizeof(u64), stream);
    return;
  }
  SwPartitionDst dst{};
  for (u32 r = 0; r < SW_MAX_RANKS; ++r) { // one send buffer: the kernel adds the per-destination prefix
    dst.xyz[r] = out_xyz;
    dst.ids[r] = out_id;
    dst.attr[r] = out_attr;
  }
  partition_count_kernel<<<tiles, PT_THREADS, 0, stream>>>(keys, n, sp, tile_counts);
  partition_scan_kernel<<<SW_MAX_RANKS, 1024, 0, stream>>>(tile_counts, tiles, send_counts);
  launch_partition_scatter(tiles, keys, xyz, out_attr ? attr : nullptr, attr_words, n, sp, tile_counts, send_counts,
                           id_base, dst, stream);
}

void
launch_partition_to_peers(const u64* keys, const double* xyz, u64 n, const u32* first_prefix, u32 n_ranks, u32 id_base,
                          u32* tile_counts, u64* send_counts, double* const* peer_xyz, u32* const* peer_ids,
                          const u64* dst_offsets, const u32* attr, u32 attr_words, u32* const* peer_attr,
                          cudaStream_t stream)
{
  const SwSplitters sp = make_splitters(first_prefix, n_ranks);
  const u32 tiles = (u32)partition_tiles(n);
  if (n == 0) {
    cudaMemsetAsync(send_counts, 0, SW_MAX_RANKS * sizeof(u64), stream);
    return;
  }
  SwPartitionDst dst{};
  for (u32 r = 0; r < n_ranks; ++r) { // this source's block inside destination r's receive buffer
    dst.xyz[r] = peer_xyz[r] + 3 * dst_offsets[r];
    dst.ids[r] = peer_ids[r] + dst_offsets[r];
    dst.attr[r] = (peer_attr && attr) ? peer_attr[r] + (u64)attr_words * dst_offsets[r] : nullptr;
  }
  partition_count_kernel<<<tiles, PT_THREADS, 0, stream>>>(keys, n, sp, tile_counts);
  partition_scan_kernel<<<SW_MAX_RANKS, 1024, 0, stream>>>(tile_counts, tiles, send_counts);
  launch_partition_scatter(tiles, keys, xyz, peer_attr ? attr : nullptr, attr_words, n, sp, tile_counts, nullptr,
                           id_base, dst, stream);
}

// ---------------------------------------------------------------------------------------------
// global node counts for nodes that span shards
// ---------------------------------------------------------------------------------------------
// dense[prefix of node r] = local point count of node r   (dense has 8^levels entries, zeroed)
__global__ void __launch_bounds__(256)
node_counts_to_dense_kernel(const u64* __restrict__ keys, const u32* __restrict__ node_start, u32 n_nodes,
                            int node_shift, u32* __restrict__ dense)
{
  const u32 r = blockIdx.x * 256 + threadIdx.x;
  if (r >= n_nodes)
    return;
  const u32 b = node_start[r];
  // saturate: the sum over <= 16 ranks must not wrap; exact for thresholds below 2^28
  const u32 c = node_start[r + 1] - b;
  dense[(keys[b] & SW_KEY_MASK) >> node_shift] = c < 0x0FFFFFFFu ? c : 0x0FFFFFFFu;
}

// gcount[r] = dense[prefix of node r]   (after the caller summed `dense` over the ranks)
__global__ void __launch_bounds__(256)
node_counts_from_dense_kernel(const u64* __restrict__ keys, const u32* __restrict__ node_start, u32 n_nodes,
                              int node_shift, const u32* __restrict__ dense, u32* __restrict__ gcount)
{
  const u32 r = blockIdx.x * 256 + threadIdx.x;
  if (r >= n_nodes)
    return;
  gcount[r] = dense[(keys[node_start[r]] & SW_KEY_MASK) >> node_shift];
}

void
launch_node_counts_to_dense(const u64* keys, const u32* node_start, u32 n_nodes, int node_shift, u32* dense,
                            cudaStream_t stream)
{
  if (n_nodes)
    node_counts_to_dense_kernel<<<(n_nodes + 255) / 256, 256, 0, stream>>>(keys, node_start, n_nodes, node_shift, dense);
}

void
launch_node_counts_from_dense(const u64* keys, const u32* node_start, u32 n_nodes, int node_shift, const u32* dense,
                              u32* gcount, cudaStream_t stream)
{
  if (n_nodes)
    node_counts_from_dense_kernel<<<(n_nodes + 255) / 256, 256, 0, stream>>>(keys, node_start, n_nodes, node_shift, dense,
                                                                          gcount);
}

// out[i] = map[perm[idx[i]]]: sorted position -> received point -> global point id
__global__ void __launch_bounds__(256)
compose_ids_mapped_kernel(const u32* __restrict__ perm, const u32* __restrict__ idx, const u32* __restrict__ map, u64 n,
                          u32* __restrict__ out)
{
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256)
    out[i] = map[perm[idx[i]]];
}

void
launch_compose_ids_mapped(const u32* perm, const u32* idx, const u32* map, u64 n, u32* out, cudaStream_t stream)
{
  if (n == 0)
    return;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const u64 want = (n + 255) / 256, cap = (u64)sms * 8;
  compose_ids_mapped_kernel<<<(u32)(want < cap ? want : cap), 256, 0, stream>>>(perm, idx, map, n, out);
}

// =============================================================================================
// MIN_DISTANCE across shard faces
// =============================================================================================
// A node above the shard prefix depth spans GPUs.  PoissonDiskSampling (Sampling.h:421-471, SparseGrid.cpp:
// 116-146) is one greedy over ALL points of the node in Morton order, so an accepted point on one GPU can
// lie within the spacing of an accepted point on the neighbouring GPU.  Running the greedy in rank order
// would serialise the GPUs; instead every GPU samples its part on its own (exact inside the part), then
//   face_flag     marks its accepted points that have another shard within the spacing,
//   face_collect  packs them (key, position) in Morton order,
//   (all-gather of those few points through the caller's hook)
//   face_resolve  rejects every own accepted point that is closer than the spacing to an accepted point of
//                 a LOWER rank (the lower rank comes first in Morton order, as in the sequential greedy).
// Rejected points move on to the child nodes like any other point that was not selected.  The merged node
// then satisfies the reference's invariant (no two stored points closer than the spacing); it can hold
// slightly fewer points than the sequential greedy would (north_star's relaxed mode: stated 1 % bound,
// checked by tests/test_gpu_sharded.py).
__global__ void __launch_bounds__(256)
face_flag_kernel(const u32* __restrict__ in_idx, const double* __restrict__ pos, const unsigned char* __restrict__ state,
                 u64 count, SwBounds b, double reach, SwSplitters sp, u32 my_rank, u32* __restrict__ flags)
{
  const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
  if (i >= count)
    return;
  u32 f = 0;
  if (state[i] == 1) {
    const u64 idx = in_idx ? in_idx[i] : i;
    const double p[3] = { pos[3 * idx], pos[3 * idx + 1], pos[3 * idx + 2] };
    // the owner cells (shard-level subtrees) are at least `reach` wide, so the 27 points p + {-r, 0, r}^3 hit
    // every owner cell the cube [p - r, p + r]^3 overlaps
    for (int dx = -1; dx <= 1 && !f; ++dx)
      for (int dy = -1; dy <= 1 && !f; ++dy)
        for (int dz = -1; dz <= 1 && !f; ++dz) {
          double q[3] = { p[0] + dx * reach, p[1] + dy * reach, p[2] + dz * reach };
#pragma unroll
          for (int a = 0; a < 3; ++a)
            q[a] = q[a] < b.min[a] ? b.min[a] : (q[a] > b.max[a] ? b.max[a] : q[a]);
          if (dest_of(morton_from_position(q[0], q[1], q[2], b), sp) != my_rank)
            f = 1;
        }
  }
  flags[i] = f;
}

__global__ void __launch_bounds__(256)
face_collect_kernel(const u64* __restrict__ in_key, const u32* __restrict__ in_idx, const double* __restrict__ pos,
                    u64 count, const u32* __restrict__ flags, const u64* __restrict__ offs, SwFaceRecord* __restrict__ rec,
                    u32* __restrict__ rec_src)
{
  const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
  if (i >= count || !flags[i])
    return;
  const u64 idx = in_idx ? in_idx[i] : i;
  SwFaceRecord r;
  r.key = in_key[i] & SW_KEY_MASK;
  r.x = pos[3 * idx];
  r.y = pos[3 * idx + 1];
  r.z = pos[3 * idx + 2];
  rec[offs[i]] = r;
  rec_src[offs[i]] = (u32)i;
}

__device__ __forceinline__ u64
face_lower_bound(const SwFaceRecord* __restrict__ a, u64 n, u64 key)
{
  u64 lo = 0, hi = n;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    if (a[mid].key < key)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128)
face_resolve_kernel(const SwFaceRecord* __restrict__ mine, const u32* __restrict__ mine_src, u32 n_mine,
                    const SwFaceRecord* __restrict__ all, SwFaceRanks fr, u32 my_rank, int cell_levels, int node_levels,
                    double threshold, unsigned char* __restrict__ state)
{
  const u32 j = blockIdx.x * 128 + threadIdx.x;
  if (j >= n_mine)
    return;
  const SwFaceRecord me = mine[j];
  const int node_shift = shift_for_levels(node_levels);
  const u64 node_lo = node_levels ? ((me.key >> node_shift) << node_shift) : 0ull;
  const u64 node_hi = node_levels ? node_lo + (1ull << node_shift) : (1ull << 63);
  const int below = cell_levels - node_levels;
  bool hit = false;
  for (u32 q = 0; q < my_rank && !hit; ++q) {
    const SwFaceRecord* list = all + fr.first[q];
    const u64 len = fr.first[q + 1] - fr.first[q];
    if (!len)
      continue;
    if (below <= 0) { // the node is not larger than a search cell: every point of the node is a candidate
      for (u64 k = face_lower_bound(list, len, node_lo); k < len && list[k].key < node_hi && !hit; ++k) {
        const double dx = me.x - list[k].x, dy = me.y - list[k].y, dz = me.z - list[k].z;
        hit = dx * dx + dy * dy + dz * dz < threshold;
      }
      continue;
    }
    const int cell_shift = shift_for_levels(cell_levels);
    const u64 code = me.key >> cell_shift;
    const long long side = 1ll << cell_levels;
    const long long x = (long long)contract_bits_by_3(code >> 2);
    const long long y = (long long)contract_bits_by_3(code >> 1);
    const long long z = (long long)contract_bits_by_3(code);
    for (int dx = -1; dx <= 1 && !hit; ++dx)
      for (int dy = -1; dy <= 1 && !hit; ++dy)
        for (int dz = -1; dz <= 1 && !hit; ++dz) {
          const long long X = x + dx, Y = y + dy, Z = z + dz;
          if (X < 0 || Y < 0 || Z < 0 || X >= side || Y >= side || Z >= side)
            continue;
          const u64 nc = expand_bits_by_3((u64)Z) | (expand_bits_by_3((u64)Y) << 1) | (expand_bits_by_3((u64)X) << 2);
          if ((nc >> (3 * below)) != (code >> (3 * below)))
            continue; // another node: the greedy is per node
          const u64 lo = nc << cell_shift, hi = cell_shift ? lo + (1ull << cell_shift) : lo + 1;
          for (u64 k = face_lower_bound(list, len, lo); k < len && list[k].key < hi && !hit; ++k) {
            const double ex = me.x - list[k].x, ey = me.y - list[k].y, ez = me.z - list[k].z;
            hit = ex * ex + ey * ey + ez * ez < threshold;
          }
        }
  }
  if (hit)
    state[mine_src[j]] = 2;
}

void
launch_face_flag(const u32* in_idx, const double* pos, const unsigned char* state, u64 count, const SwBounds& b,
                 double reach, const u32* first_prefix, u32 n_ranks, u32 my_rank, u32* flags, cudaStream_t stream)
{
  if (!count)
    return;
  const SwSplitters sp = make_splitters(first_prefix, n_ranks);
  face_flag_kernel<<<(u32)((count + 255) / 256), 256, 0, stream>>>(in_idx, pos, state, count, b, reach, sp, my_rank,
                                                                  flags);
}

void
launch_face_collect(const u64* in_key, const u32* in_idx, const double* pos, u64 count, const u32* flags,
                    const u64* offs, SwFaceRecord* rec, u32* rec_src, cudaStream_t stream)
{
  if (count)
    face_collect_kernel<<<(u32)((count + 255) / 256), 256, 0, stream>>>(in_key, in_idx, pos, count, flags, offs, rec,
                                                                       rec_src);
}

void
launch_face_resolve(const SwFaceRecord* mine, const u32* mine_src, u32 n_mine, const SwFaceRecord* all,
                    const SwFaceRanks& fr, u32 my_rank, int cell_levels, int node_levels, double threshold,
                    unsigned char* state, cudaStream_t stream)
{
  if (n_mine && my_rank)
    face_resolve_kernel<<<(n_mine + 127) / 128, 128, 0, stream>>>(mine, mine_src, n_mine, all, fr, my_rank,
                                                                  cell_levels, node_levels, threshold, state);
}
