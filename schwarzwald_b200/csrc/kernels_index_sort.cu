// kernels_index_sort.cu — K1 (Morton indexing), K2 (onesweep LSD radix sort), K4 (gathers).
//
// Reference behaviour replaced (paths relative to /root/reference/schwarzwald/core):
//   K1  index_point<21> / calculate_morton_index<21>   tiling/OctreeAlgorithms.h:64-87,145-175
//       (driven by parallel::transform / scatter, tiling/TilingAlgorithms.cpp:588-598,1262-1285)
//   K2  std::sort over IndexedPoint64 by key           tiling/TilingAlgorithms.cpp:600-604,1289-1292
//       tie rule: stable w.r.t. the original point index (SURVEY.md §8a "S")
//   K4  the gather that persist_points performs through PointReference
//       (tiling/TilingAlgorithms.cpp:232-236,330-334)
//
// HBM layout: positions stay AoS (x,y,z doubles, 24 B) exactly as PointBuffer holds them
// (datastructures/PointBuffer.h:291); keys are one u64 per point; the sort payload is the u32
// original index.  All kernels are bandwidth bound; no tensor-core work exists on this path.
#include "swgpu_internal.cuh"

// =============================================================================================
// K1  Morton encode (+ digit histograms for the sort)
// =============================================================================================
#define MORTON_THREADS 256

// Radix of the sort (K2).  A MortonIndex64 has 63 significant bits: 9-bit digits sort it in 7 passes
// instead of the 8 that 8-bit digits need (one read + one write of every pair less).  One thread per
// digit value in K2, so a CTA has RS_RADIX threads.
#ifndef RS_BITS
#define RS_BITS 8
#endif
#define RS_RADIX (1 << RS_BITS)
#define RS_PASSES ((63 + RS_BITS - 1) / RS_BITS)
#define RS_DIGIT_MASK ((u32)(RS_RADIX - 1))

// std::min(bmax, std::max(bmin, p)) with the exact comparison order of libstdc++
__device__ __forceinline__ double
clamp_like_reference(double p, double bmin, double bmax)
{
  const double mx = (bmin < p) ? p : bmin;
  return (mx < bmax) ? mx : bmax;
}

__device__ __forceinline__ bool
index_one(double& x, double& y, double& z, const SwBounds& b)
{
  // AABB::isInside is inclusive on both ends (math/AABB.h:27-31)
  const bool inside = x >= b.min[0] && x <= b.max[0] && y >= b.min[1] && y <= b.max[1] && z >= b.min[2] &&
                      z <= b.max[2];
  if (!inside) {
    x = clamp_like_reference(x, b.min[0], b.max[0]);
    y = clamp_like_reference(y, b.min[1], b.max[1]);
    z = clamp_like_reference(z, b.min[2], b.max[2]);
  }
  return !inside;
}

__device__ __forceinline__ void
hist_add(u32* s_hist, u64 key)
{
#pragma unroll
  for (int p = 0; p < RS_PASSES; ++p)
    atomicAdd(&s_hist[p * RS_RADIX + ((u32)(key >> (RS_BITS * p)) & RS_DIGIT_MASK)], 1u);
}

// Each thread indexes two consecutive points: 48 bytes = three 16-byte loads.
__global__ void __launch_bounds__(MORTON_THREADS)
morton_encode_kernel(double* __restrict__ xyz, u64 n, SwBounds b, u64* __restrict__ keys, u32* __restrict__ hist,
                     u32* __restrict__ n_clamped)
{
  __shared__ u32 s_hist[RS_PASSES * RS_RADIX];
  for (int i = threadIdx.x; i < RS_PASSES * RS_RADIX; i += MORTON_THREADS)
    s_hist[i] = 0;
  __syncthreads();

  const u64 n_pairs = n >> 1;
  u32 clamped = 0;
  for (u64 pair = (u64)blockIdx.x * MORTON_THREADS + threadIdx.x; pair < n_pairs;
       pair += (u64)gridDim.x * MORTON_THREADS) {
    double2* p2 = reinterpret_cast<double2*>(xyz) + 3 * pair;
    double2 a = p2[0], c = p2[1], e = p2[2];
    double x0 = a.x, y0 = a.y, z0 = c.x, x1 = c.y, y1 = e.x, z1 = e.y;
    const bool c0 = index_one(x0, y0, z0, b);
    const bool c1 = index_one(x1, y1, z1, b);
    if (c0 | c1) { // write the clamped coordinates back (OctreeAlgorithms.h:167-169)
      p2[0] = make_double2(x0, y0);
      p2[1] = make_double2(z0, x1);
      p2[2] = make_double2(y1, z1);
      clamped += (u32)c0 + (u32)c1;
    }
    const u64 k0 = morton_from_position(x0, y0, z0, b);
    const u64 k1 = morton_from_position(x1, y1, z1, b);
    reinterpret_cast<ulonglong2*>(keys)[pair] = make_ulonglong2(k0, k1);
    hist_add(s_hist, k0);
    hist_add(s_hist, k1);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double* p = xyz + 3 * (n - 1);
    double x = p[0], y = p[1], z = p[2];
    if (index_one(x, y, z, b)) {
      p[0] = x;
      p[1] = y;
      p[2] = z;
      ++clamped;
    }
    const u64 k = morton_from_position(x, y, z, b);
    keys[n - 1] = k;
    hist_add(s_hist, k);
  }
  if (clamped)
    atomicAdd(n_clamped, clamped);
  __syncthreads();
  for (int i = threadIdx.x; i < RS_PASSES * RS_RADIX; i += MORTON_THREADS) {
    const u32 v = s_hist[i];
    if (v)
      atomicAdd(&hist[i], v);
  }
}

__global__ void __launch_bounds__(MORTON_THREADS)
key_histogram_kernel(const u64* __restrict__ keys, u64 n, u32* __restrict__ hist)
{
  __shared__ u32 s_hist[RS_PASSES * RS_RADIX];
  for (int i = threadIdx.x; i < RS_PASSES * RS_RADIX; i += MORTON_THREADS)
    s_hist[i] = 0;
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * MORTON_THREADS + threadIdx.x; i < n; i += (u64)gridDim.x * MORTON_THREADS)
    hist_add(s_hist, keys[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < RS_PASSES * RS_RADIX; i += MORTON_THREADS) {
    const u32 v = s_hist[i];
    if (v)
      atomicAdd(&hist[i], v);
  }
}

static int
persistent_grid(u64 work_items, int threads, int ctas_per_sm)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  u64 want = (work_items + threads - 1) / threads;
  u64 cap = (u64)sms * ctas_per_sm;
  if (want < 1)
    want = 1;
  return (int)(want < cap ? want : cap);
}

void
launch_morton_encode(double* xyz, u64 n, const SwBounds& b, u64* keys, u32* hist, u32* n_clamped, cudaStream_t stream)
{
  if (n == 0)
    return;
  const int grid = persistent_grid((n + 1) / 2, MORTON_THREADS, 8);
  morton_encode_kernel<<<grid, MORTON_THREADS, 0, stream>>>(xyz, n, b, keys, hist, n_clamped);
}

void
launch_key_histogram(const u64* keys, u64 n, u32* hist, cudaStream_t stream)
{
  if (n == 0)
    return;
  const int grid = persistent_grid(n, MORTON_THREADS, 8);
  key_histogram_kernel<<<grid, MORTON_THREADS, 0, stream>>>(keys, n, hist);
}

// ---------------------------------------------------------------------------------------------
// K1-LAS    LAS record coordinates -> PointBuffer position -> index_point
//   position_from_las_point              io/LASFile.cpp:79-94     offset + X * scale (two roundings),
//                                                                 clamped into the LAS header bounds
//   the tiler's point transformation     process/TilerProcess.cpp:552-559  shift to the centre of the
//                                                                 cubic bounds, round to float32
//   index_point<21>                      as K1
// One host pass over every point (and half of the PCIe bytes: 12 instead of 24 per point) disappears;
// the doubles the reference would hold in its PointBuffer are written to xyz_out for the sampling
// kernels and the writers.  Each thread converts four records = three 16-byte loads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double
las_axis(int v, int a, const SwLasTransform& t)
{
  double p = t.offset[a] + (double)v * t.scale[a];
  p = clamp_like_reference(p, t.hmin[a], t.hmax[a]);
  if (t.shift) {
    p = p - t.center[a];
    p = (double)(float)p;
  }
  return p;
}

__device__ __forceinline__ u64
las_one(int X, int Y, int Z, const SwLasTransform& t, const SwBounds& b, double* out, u32& clamped)
{
  double x = las_axis(X, 0, t), y = las_axis(Y, 1, t), z = las_axis(Z, 2, t);
  clamped += (u32)index_one(x, y, z, b);
  out[0] = x;
  out[1] = y;
  out[2] = z;
  return morton_from_position(x, y, z, b);
}

__global__ void __launch_bounds__(MORTON_THREADS)
las_encode_kernel(const int* __restrict__ las, u64 n, SwLasTransform t, SwBounds b, double* __restrict__ xyz_out,
                  u64* __restrict__ keys, u32* __restrict__ hist, u32* __restrict__ n_clamped)
{
  __shared__ u32 s_hist[RS_PASSES * RS_RADIX];
  for (int i = threadIdx.x; i < RS_PASSES * RS_RADIX; i += MORTON_THREADS)
    s_hist[i] = 0;
  __syncthreads();

  const u64 n_quads = n >> 2;
  u32 clamped = 0;
  for (u64 q = (u64)blockIdx.x * MORTON_THREADS + threadIdx.x; q < n_quads; q += (u64)gridDim.x * MORTON_THREADS) {
    const int4* p4 = reinterpret_cast<const int4*>(las) + 3 * q;
    const int4 a = p4[0], c = p4[1], e = p4[2];
    double o[12];
    u64 k[4];
    k[0] = las_one(a.x, a.y, a.z, t, b, o + 0, clamped);
    k[1] = las_one(a.w, c.x, c.y, t, b, o + 3, clamped);
    k[2] = las_one(c.z, c.w, e.x, t, b, o + 6, clamped);
    k[3] = las_one(e.y, e.z, e.w, t, b, o + 9, clamped);
    double2* d2 = reinterpret_cast<double2*>(xyz_out) + 6 * q;
#pragma unroll
    for (int i = 0; i < 6; ++i)
      d2[i] = make_double2(o[2 * i], o[2 * i + 1]);
    ulonglong2* k2 = reinterpret_cast<ulonglong2*>(keys) + 2 * q;
    k2[0] = make_ulonglong2(k[0], k[1]);
    k2[1] = make_ulonglong2(k[2], k[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      hist_add(s_hist, k[i]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (u32)(n & 3)) { // ragged tail: up to three records
    const u64 i = (n & ~3ull) + threadIdx.x;
    double o[3];
    const u64 k = las_one(las[3 * i], las[3 * i + 1], las[3 * i + 2], t, b, o, clamped);
    xyz_out[3 * i] = o[0];
    xyz_out[3 * i + 1] = o[1];
    xyz_out[3 * i + 2] = o[2];
    keys[i] = k;
    hist_add(s_hist, k);
  }
  if (clamped)
    atomicAdd(n_clamped, clamped);
  __syncthreads();
  for (int i = threadIdx.x; i < RS_PASSES * RS_RADIX; i += MORTON_THREADS) {
    const u32 v = s_hist[i];
    if (v)
      atomicAdd(&hist[i], v);
  }
}

void
launch_las_encode(const int* las, u64 n, const SwLasTransform& t, const SwBounds& b, double* xyz_out, u64* keys,
                  u32* hist, u32* n_clamped, cudaStream_t stream)
{
  if (n == 0)
    return;
  const int grid = persistent_grid((n + 3) / 4, MORTON_THREADS, 8);
  las_encode_kernel<<<grid, MORTON_THREADS, 0, stream>>>(las, n, t, b, xyz_out, keys, hist, n_clamped);
}

// =============================================================================================
// K2  onesweep LSD radix sort, 8-bit digits, u64 keys + u32 payload
// =============================================================================================
#define RS_THREADS RS_RADIX
#define RS_WARPS (RS_THREADS / 32)
#ifndef RS_ITEMS
#define RS_ITEMS 16
#endif
#define RS_TILE (RS_THREADS * RS_ITEMS) // pairs per tile: 4096 (8-bit digits) or 8192 (9-bit)
#ifndef RS_MATCH_MODE
#define RS_MATCH_MODE 2 /* 0 = __match_any_sync, 1 = eight ballots, 2 = shared-memory atomicOr */
#endif
#ifndef RS_LOOK
#define RS_LOOK 4 /* look-back descriptors fetched per round trip (measured: 1: 6.21, 2: 5.87, 4: 5.79, 8: 5.87, 16: 6.12 ms per sort) */
#endif
#ifndef RS_MIN_CTAS
#define RS_MIN_CTAS (1024 / RS_THREADS) /* CTAs per SM the register allocation is capped for: 32 warps per SM */
#endif

// RS_EXPERIMENT: bit mask of measurement-only switches that remove one phase of the pass to see what it
// costs (tools/bench_sort.py; the output is NOT sorted when any bit is set, never set in the product build):
//   1 = no look-back (global offsets = digit base only)   2 = no ranking (trivial ranks)
//   4 = no scatter (coalesced tile-order stores)           8 = no early-count atomics
#ifndef RS_EXPERIMENT
#define RS_EXPERIMENT 0
#endif
#ifndef RS_SPLIT_TABLE
#define RS_SPLIT_TABLE 1 /* 0 = {mask, count} entries of 8 bytes, 1 = separate arrays (5 % faster: 32 banks instead of 16 bank pairs), 2 = + alternating mask arrays (no further gain), 3 = two rows per round (slower: more table reads) */
#endif

#define RS_FLAG_AGG (1u << 30)
#define RS_FLAG_PFX (2u << 30)
#define RS_VAL_MASK ((1u << 30) - 1)

// exclusive scan of the RS_PASSES x RS_RADIX histogram rows, in place (one block, one row per warp)
__global__ void
digit_base_kernel(u32* __restrict__ hist)
{
  constexpr int PER_LANE = RS_RADIX / 32;
  const int row = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  u32* h = hist + row * RS_RADIX;
  u32 v[PER_LANE];
  u32 sum = 0;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    v[i] = h[lane * PER_LANE + i];
    sum += v[i];
  }
  u32 incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o)
      incl += t;
  }
  u32 run = incl - sum;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    h[lane * PER_LANE + i] = run;
    run += v[i];
  }
}

// One pass = one read and one write of every (key, id) pair.  Per tile of 4096 pairs:
//   1. early counts   per-warp digit histograms by shared-memory atomics; the tile's digit counts
//                     are published (AGG) before any ranking work, so successors never spin on us
//   2. ranking        stable rank of every key inside the tile: __match_any_sync groups the lanes
//                     holding the same digit, the group leader reserves the group's slots in the
//                     warp's running offset, lanes keep their lane order
//   3. exchange       keys (then ids) go through shared memory in ranked order so that the global
//                     writes are runs of consecutive addresses per digit
//   4. look-back      thread d resolves digit d's global offset (decoupled look-back) between the
//                     key and the id exchange, long after step 1 published this tile's counts
// The warp histograms alias the exchange buffer (they are dead once the ranks are final), which
// keeps a CTA at ~50 KB of shared memory: four CTAs per SM.
template<int PASS, bool FIRST>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_CTAS)
onesweep_pass_kernel(const u64* __restrict__ keys_in, const u32* __restrict__ vals_in, u64* __restrict__ keys_out,
                     u32* __restrict__ vals_out, u32 n, const u32* __restrict__ digit_base, u32* __restrict__ status,
                     u32* __restrict__ ticket)
{
  constexpr int SHIFT = RS_BITS * PASS;
  extern __shared__ __align__(16) unsigned char smem[];
  u64* s_keys = reinterpret_cast<u64*>(smem);               // RS_TILE * 8
  u32* s_vals = reinterpret_cast<u32*>(smem + RS_TILE * 8); // RS_TILE * 4
  // per (warp, digit) table entry: .x = lanes currently holding the digit (RS_MATCH_MODE 2),
  // .y = count, later the next free tile-local rank.  Aliases s_keys.
  uint2* s_tab = reinterpret_cast<uint2*>(smem);             // RS_WARPS * 256 * 8 B
#if RS_SPLIT_TABLE
  // the same table as two u32 arrays: an 8-byte entry spans two banks, so 256 entries only spread over 16
  // bank pairs; separate mask / count arrays spread the warp's 32 accesses over all 32 banks
  u32* s_cnt = reinterpret_cast<u32*>(smem);                 // RS_WARPS * RS_RADIX
  u32* s_msk = s_cnt + RS_WARPS * RS_RADIX;                  // RS_WARPS * RS_RADIX (x 2 when RS_SPLIT_TABLE == 2)
#endif
  u32* s_gofs = reinterpret_cast<u32*>(smem + RS_TILE * 12); // 256: global base - tile-local offset
  u32* s_wsum = s_gofs + RS_RADIX;                           // RS_WARPS
  __shared__ u32 s_tile;

  const u32 tid = threadIdx.x;
  const u32 warp = tid >> 5;
  const u32 lane = tid & 31;

  if (tid == 0)
    s_tile = atomicAdd(ticket, 1u);
#pragma unroll
  for (int i = 0; i < RS_WARPS; ++i)
    s_tab[i * RS_RADIX + tid] = make_uint2(0u, 0u);
#if RS_SPLIT_TABLE >= 2
#pragma unroll
  for (int i = 0; i < RS_WARPS; ++i)
    s_msk[(RS_WARPS + i) * RS_RADIX + tid] = 0u; // the second mask array
#endif
  __syncthreads();
  const u32 tile = s_tile;
  const u32 tile_base = tile * RS_TILE;
  const u32 valid = (n - tile_base) < RS_TILE ? (n - tile_base) : RS_TILE;
  const bool full = valid == RS_TILE;

  // ---- load (warp-striped: item j of a warp is 32 consecutive pairs) -------------------------
  u64 key[RS_ITEMS];
  const u32 warp_base = warp * (32 * RS_ITEMS) + lane;
  if (full) {
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j)
      key[j] = keys_in[tile_base + warp_base + j * 32];
  } else {
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      const u32 p = warp_base + j * 32;
      key[j] = (p < valid) ? keys_in[tile_base + p] : ~0ull;
    }
  }

  // ---- 1. early counts --------------------------------------------------------------------------
  uint2* my_tab = s_tab + warp * RS_RADIX;
  (void)my_tab;
#if RS_SPLIT_TABLE
  u32* my_cnt = s_cnt + warp * RS_RADIX;
  u32* my_msk = s_msk + warp * RS_RADIX;
#endif
#if !(RS_EXPERIMENT & 8)
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j)
#if RS_SPLIT_TABLE
    atomicAdd(&my_cnt[(u32)(key[j] >> SHIFT) & RS_DIGIT_MASK], 1u);
#else
    atomicAdd(&my_tab[(u32)(key[j] >> SHIFT) & RS_DIGIT_MASK].y, 1u);
#endif
#endif
  __syncthreads();

  u32 my_excl;  // tile-local exclusive offset of digit `tid`
  u32 my_count; // keys of the tile with digit `tid`
  {
    const u32 d = tid; // RS_THREADS == RS_RADIX
    u32 c[RS_WARPS];
    u32 run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
#if RS_SPLIT_TABLE
      c[w] = s_cnt[w * RS_RADIX + d];
#else
      c[w] = s_tab[w * RS_RADIX + d].y;
#endif
      run += c[w];
    }
    my_count = run;
    u32* my_status = status + (size_t)tile * RS_RADIX + d;
    st_relaxed_u32(my_status, (tile == 0 ? RS_FLAG_PFX : RS_FLAG_AGG) | run);
    // exclusive scan of the tile counts over the 256 digits
    u32 incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o)
        incl += t;
    }
    if (lane == 31)
      s_wsum[warp] = incl;
    __syncthreads();
    u32 wofs = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w)
      wofs += (w < (int)warp) ? s_wsum[w] : 0u;
    my_excl = wofs + incl - run;
    // warp counts -> first tile-local rank of (warp, digit)
    u32 acc = my_excl;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
#if RS_SPLIT_TABLE
      s_cnt[w * RS_RADIX + d] = acc;
#else
      s_tab[w * RS_RADIX + d].y = acc;
#endif
      acc += c[w];
    }
  }
  __syncthreads();

  // ---- 2. ranking ---------------------------------------------------------------------------------
  // Lanes holding the same digit form a group; the group takes the next `size` ranks of its
  // (warp, digit) entry in lane order.  RS_MATCH_MODE picks how the group is found: measured on
  // B200, __match_any_sync costs ~60 SM-cycles per warp instruction (ADU pipe), eight ballots ~25,
  // a shared-memory atomicOr + load ~7.
  unsigned short rank[RS_ITEMS];
  {
    const u32 lt = lanemask_lt();
#if RS_EXPERIMENT & 2
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j)
      rank[j] = (unsigned short)(warp_base - lane + j * 32 + lane);
    (void)lt;
#elif RS_MATCH_MODE == 2 && RS_SPLIT_TABLE == 3
    // two rows per round: row j marks its lanes in mask array A, row j + 1 in mask array B; one set of
    // warp barriers serves both rows.  A digit's running count is advanced once per round by exactly one
    // lane: the leader of its row-(j+1) group if there is one, else the leader of its row-j group.
    const u32 lane_bit = 1u << lane;
    u32* mskA = my_msk;
    u32* mskB = my_msk + RS_WARPS * RS_RADIX;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j += 2) {
      const u32 d0 = (u32)(key[j] >> SHIFT) & RS_DIGIT_MASK;
      const u32 d1 = (u32)(key[j + 1] >> SHIFT) & RS_DIGIT_MASK;
      atomicOr(&mskA[d0], lane_bit);
      atomicOr(&mskB[d1], lane_bit);
      __syncwarp();
      const u32 a0 = mskA[d0], b0 = mskB[d0], c0 = my_cnt[d0];
      const u32 a1 = mskA[d1], b1 = mskB[d1], c1 = my_cnt[d1];
      __syncwarp();
      rank[j] = (unsigned short)(c0 + __popc(a0 & lt));
      rank[j + 1] = (unsigned short)(c1 + __popc(a1) + __popc(b1 & lt));
      if ((a0 >> lane) <= 1u && b0 == 0u) { // leader of a row-j group whose digit does not occur in row j + 1
        my_cnt[d0] = c0 + __popc(a0);
        mskA[d0] = 0u;
      }
      if ((b1 >> lane) <= 1u) { // leader of a row-(j+1) group: accounts for the row-j group of the digit too
        my_cnt[d1] = c1 + __popc(a1) + __popc(b1);
        mskB[d1] = 0u;
        if (a1)
          mskA[d1] = 0u;
      }
      __syncwarp();
    }
#elif RS_MATCH_MODE == 2
    const u32 lane_bit = 1u << lane;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
#if RS_SPLIT_TABLE
      const u32 dg = (u32)(key[j] >> SHIFT) & RS_DIGIT_MASK;
#if RS_SPLIT_TABLE == 2
      // rows alternate between two mask arrays: the leader's clear of row j cannot race with the ORs of
      // row j + 1, which saves the third warp barrier of every row
      u32* msk = my_msk + (j & 1) * (RS_WARPS * RS_RADIX);
#else
      u32* msk = my_msk;
#endif
      atomicOr(&msk[dg], lane_bit);
      __syncwarp();
      uint2 v;
      v.x = msk[dg];    // group
      v.y = my_cnt[dg]; // first free rank
      __syncwarp();
      const u32 lower = __popc(v.x & lt);
      if ((v.x >> lane) <= 1u) { // highest lane of the group: reserve the ranks, clear the group
        msk[dg] = 0u;
        my_cnt[dg] = v.y + lower + 1u;
      }
#if RS_SPLIT_TABLE != 2
      __syncwarp();
#endif
      rank[j] = (unsigned short)(v.y + lower);
#else
      uint2* e = &my_tab[(u32)(key[j] >> SHIFT) & RS_DIGIT_MASK];
      atomicOr(&e->x, lane_bit);
      __syncwarp();
      const uint2 v = *e; // .x = group, .y = first free rank
      __syncwarp();
      const u32 lower = __popc(v.x & lt);
      if ((v.x >> lane) <= 1u) // highest lane of the group: reserve the ranks, clear the group
        *e = make_uint2(0u, v.y + lower + 1u);
      __syncwarp();
      rank[j] = (unsigned short)(v.y + lower);
#endif
    }
#else
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      const u32 d = (u32)(key[j] >> SHIFT) & RS_DIGIT_MASK;
#if RS_MATCH_MODE == 1
      u32 peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const u32 vote = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? vote : ~vote;
      }
#else
      const u32 peers = __match_any_sync(0xffffffffu, d);
#endif
      const u32 lower = __popc(peers & lt);
#if RS_SPLIT_TABLE
      const u32 base = my_cnt[d];
      __syncwarp();
      if ((peers >> lane) <= 1u)
        my_cnt[d] = base + lower + 1u;
#else
      const u32 base = my_tab[d].y;
      __syncwarp();
      if ((peers >> lane) <= 1u)
        my_tab[d].y = base + lower + 1u;
#endif
      __syncwarp();
      rank[j] = (unsigned short)(base + lower);
    }
#endif
  }
  __syncthreads(); // ranks are final: the histograms may be overwritten

  // ---- 3a. keys through shared memory --------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j)
    s_keys[rank[j]] = key[j];

  // ids are loaded only now: their latency overlaps the look-back below
  u32 val[RS_ITEMS];
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const u32 p = warp_base + j * 32;
    if (FIRST)
      val[j] = tile_base + p;
    else
      val[j] = (full || p < valid) ? vals_in[tile_base + p] : 0u;
  }

  // ---- 4. look-back for digit `tid` ---------------------------------------------------------------
  {
    const u32 d = tid;
    u32 prev = 0;
    if (tile != 0 && !(RS_EXPERIMENT & 1)) {
      u32* my_status = status + (size_t)tile * RS_RADIX + d;
      // RS_LOOK predecessors are fetched per round trip (their descriptors were published before
      // their ranking started, so they are almost always present): ~9 dependent L2 latencies
      // become ~2
      int t = (int)tile - 1;
      bool done = false;
      while (!done) {
        u32 sv[RS_LOOK];
#pragma unroll
        for (int k = 0; k < RS_LOOK; ++k)
          sv[k] = (t - k >= 0) ? ld_relaxed_u32(status + (size_t)(t - k) * RS_RADIX + d) : RS_FLAG_PFX;
#pragma unroll
        for (int k = 0; k < RS_LOOK; ++k) {
          if (!done) {
            u32 v = sv[k];
            while ((v >> 30) == 0)
              v = ld_relaxed_u32(status + (size_t)(t - k) * RS_RADIX + d);
            prev += v & RS_VAL_MASK;
            done = (v >> 30) == 2;
          }
        }
        t -= RS_LOOK;
      }
      st_relaxed_u32(my_status, RS_FLAG_PFX | ((prev + my_count) & RS_VAL_MASK));
    }
#if RS_EXPERIMENT & 1
    prev = tile * my_count; // keeps the write pattern spread out (roughly where the real offsets are)
#endif
    s_gofs[d] = digit_base[d] + prev - my_excl;
  }

  // ---- 3b. ids through shared memory ----------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j)
    s_vals[rank[j]] = val[j];
  __syncthreads();

  if (full) {
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
      const u32 p = tid + k * RS_THREADS;
      const u64 kk = s_keys[p];
#if RS_EXPERIMENT & 4
      const u32 dst = tile_base + p;
#elif RS_EXPERIMENT & 1
      u32 dst = s_gofs[(u32)(kk >> SHIFT) & RS_DIGIT_MASK] + p;
      dst = dst < n ? dst : dst % n;
#else
      const u32 dst = s_gofs[(u32)(kk >> SHIFT) & RS_DIGIT_MASK] + p;
#endif
      keys_out[dst] = kk;
      vals_out[dst] = s_vals[p];
    }
  } else {
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
      const u32 p = tid + k * RS_THREADS;
      if (p < valid) {
        const u64 kk = s_keys[p];
        u32 dst = s_gofs[(u32)(kk >> SHIFT) & RS_DIGIT_MASK] + p;
#if RS_EXPERIMENT
        dst = (RS_EXPERIMENT & 4) ? tile_base + p : (dst < n ? dst : dst % n);
#endif
        keys_out[dst] = kk;
        vals_out[dst] = s_vals[p];
      }
    }
  }
}

#define RS_SMEM_BYTES (RS_TILE * 12 + (RS_RADIX + RS_WARPS) * 4)

size_t
sort_status_words(u64 n)
{
  const size_t tiles = (size_t)((n + RS_TILE - 1) / RS_TILE);
  return (tiles ? tiles : 1) * RS_RADIX;
}

template<int PASS, bool FIRST>
static void
launch_onesweep_pass(const u64* kin, const u32* vin, u64* kout, u32* vout, u32 n, const u32* digit_base, u32* status,
                     u32* ticket, u32 tiles, cudaStream_t stream)
{
  auto kernel = onesweep_pass_kernel<PASS, FIRST>;
  // function attributes are per device: one flag per device ordinal (a process may drive several GPUs)
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM_BYTES);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (dev >= 0 && dev < 64)
      attr_set[dev] = true;
  }
  kernel<<<tiles, RS_THREADS, RS_SMEM_BYTES, stream>>>(kin, vin, kout, vout, n, digit_base, status, ticket);
}

size_t
sort_hist_words()
{
  return (size_t)RS_PASSES * RS_RADIX;
}

int
sort_passes()
{
  return RS_PASSES;
}

int
sort_input_buffer()
{
  return RS_PASSES & 1; // an odd number of ping-pongs ends in buffer 0 when it starts in buffer 1
}

int
sort_input_buffer_top(int first_pass)
{
  return (RS_PASSES - first_pass) & 1;
}

// passes first_pass .. RS_PASSES - 1, ping-ponging from (kin, vin); `implicit_ids`: the first of them generates
// the ids 0..n-1 instead of reading vin
static void
run_passes(u64* kin, u32* vin, u64* kout, u32* vout, u32 n, int first_pass, bool implicit_ids, const u32* hist,
           u32* status, u32* ticket, cudaStream_t stream)
{
  const u32 tiles = (u32)((n + RS_TILE - 1) / RS_TILE);
  cudaMemsetAsync(ticket, 0, 8 * sizeof(u32), stream);
  for (int pass = first_pass; pass < RS_PASSES; ++pass) {
    cudaMemsetAsync(status, 0, (size_t)tiles * RS_RADIX * sizeof(u32), stream);
    const u32* base = hist + pass * RS_RADIX;
    u32* tk = ticket + pass;
    const bool first = implicit_ids && pass == first_pass;
#define RS_LAUNCH(P)                                                                                                   \
  case P:                                                                                                              \
    if (first)                                                                                                         \
      launch_onesweep_pass<P, true>(kin, vin, kout, vout, n, base, status, tk, tiles, stream);                         \
    else                                                                                                               \
      launch_onesweep_pass<P, false>(kin, vin, kout, vout, n, base, status, tk, tiles, stream);                        \
    break;
    switch (pass) {
      RS_LAUNCH(0)
      RS_LAUNCH(1)
      RS_LAUNCH(2)
      RS_LAUNCH(3)
      RS_LAUNCH(4)
      RS_LAUNCH(5)
#if RS_PASSES > 7
      RS_LAUNCH(6)
      default:
        if (first)
          launch_onesweep_pass<7, true>(kin, vin, kout, vout, n, base, status, tk, tiles, stream);
        else
          launch_onesweep_pass<7, false>(kin, vin, kout, vout, n, base, status, tk, tiles, stream);
        break;
#else
      default:
        if (first)
          launch_onesweep_pass<6, true>(kin, vin, kout, vout, n, base, status, tk, tiles, stream);
        else
          launch_onesweep_pass<6, false>(kin, vin, kout, vout, n, base, status, tk, tiles, stream);
        break;
#endif
    }
#undef RS_LAUNCH
    u64* tk2 = kin;
    kin = kout;
    kout = tk2;
    u32* tv = vin;
    vin = vout;
    vout = tv;
  }
}

void
launch_radix_sort(u64* keys0, u64* keys1, u32* vals0, u32* vals1, u64 n, u32* hist, u32* status, u32* ticket,
                  cudaStream_t stream)
{
  if (n == 0)
    return;
  digit_base_kernel<<<1, RS_PASSES * 32, 0, stream>>>(hist);
  const bool from1 = sort_input_buffer() == 1;
  run_passes(from1 ? keys1 : keys0, from1 ? vals1 : vals0, from1 ? keys0 : keys1, from1 ? vals0 : vals1, (u32)n, 0,
             true, hist, status, ticket, stream);
}

// ---- top-digit sort + segment finish ----------------------------------------------------------------------
// An LSD sort has to move every pair once per digit, whatever the data.  A MortonIndex64 of a real cloud is
// nearly unique long before its last bit: 100 M terrain points fall into 62 M cells of octree level 13, the
// longest run of points sharing the top 39 key bits is 12.  So only the TOP digits are sorted by onesweep passes
// (passes first_pass .. 7: stable, ids ascending inside every run of equal top bits), and one kernel finishes
// the runs ("segments") in place: every element counts the elements of its segment that have to precede it
// (smaller low bits, or equal low bits and an earlier position = smaller id) and moves there.  That is one read
// of the keys and the ids and a write of the elements that move, instead of first_pass read+write passes.
//
// A segment is finished by the tile in which it STARTS; the tile keeps a window of FIN_LIMIT further keys in
// shared memory for segments that reach into the next tile, so segments of up to FIN_LIMIT elements are
// handled and tiles never write to the same elements (other tiles only compare the top bits of foreign
// elements, which a permutation inside a segment does not change).  A longer segment is left as it is; if it is
// not already in order (identical points are) an element that sees an inversion raises stats[0] and the caller
// runs the eight LSD passes over the current arrangement (still stable: equal keys are in id order).
#define FIN_THREADS 256
#define FIN_TILE 4096
#define FIN_LIMIT 256
#define FIN_WINDOW (FIN_TILE + FIN_LIMIT)
#define FIN_ITEMS (FIN_WINDOW / FIN_THREADS)
#define FIN_HEAD 0x80000000u    /* the element starts a run of equal top bits */
#define FIN_FOREIGN 0x40000000u /* left sentinel: the run continues from the previous tile */
#define FIN_KEEP 0xffffu
#define FIN_UNSORTED 0x80000000u /* stats[0]: a run longer than FIN_LIMIT is not in order */
#define FIN_STRIDE 24 /* elements a warp advances per 32-lane window */

// Shared memory holds one 32-bit tag per window element (the low key bits, at most 24, and the head flag: a step
// of a scan is one LDS and three integer instructions) and, per window position, which element moves there.
// Elements are pulled: the thread of position q reads the id of the element that belongs at q, and after a block
// barrier rewrites the low bits of keys[q] (the top bits are those of its run) and ids[q].  So every thread writes
// its own position only, and all reads of ids precede all writes.
// One element of a run, ranked by scanning the tags to both ends of the run (at most FIN_LIMIT - 1 steps).
template<int LOW_BITS>
__device__ __forceinline__ void
finish_scan_run(u32 j, u32 tag, const u32* s_tag, unsigned short* s_src, u32& steps, u32& moved, bool& unsorted_long,
                u32& long_elements, u32 base, u32* __restrict__ stats, u32* __restrict__ long_runs)
{
  constexpr u32 LOW_MASK = (1u << LOW_BITS) - 1;
  const u32 lo = tag & LOW_MASK;
  u32 budget = FIN_LIMIT - 1; // other elements a run may have
  bool too_long = false;
  // elements of the run in front of this one: those that do not have larger low bits stay in front
  int l = (int)j;
  u32 t = tag, rk = 0;
  while (!(t & FIN_HEAD)) {
    if (budget == 0) {
      too_long = true;
      break;
    }
    --budget;
    --l;
    t = s_tag[1 + l];
    rk += ((t & LOW_MASK) <= lo) ? 1u : 0u;
  }
  bool skip = false;
  if (!too_long) {
    // runs that start in an earlier tile are finished there (its window reaches this element), runs that start
    // behind this tile by the next one
    skip = (t & FIN_FOREIGN) || l >= FIN_TILE;
    // elements of the run behind this one: those with smaller low bits move in front
    for (u32 r = j + 1; !skip; ++r) {
      const u32 t2 = s_tag[1 + r];
      if (t2 & FIN_HEAD)
        break;
      if (budget == 0) {
        too_long = true;
        break;
      }
      --budget;
      rk += ((t2 & LOW_MASK) < lo) ? 1u : 0u;
    }
  }
  steps += FIN_LIMIT - 1 - budget;
  if (too_long) { // the run stays as it is: fine if it is in order already (identical points are)
    unsorted_long |= !(tag & FIN_HEAD) && (s_tag[j] & LOW_MASK) > lo;
    long_elements += j < FIN_TILE ? 1u : 0u;
    if (tag & FIN_HEAD) // its first element (seen by the tile that owns the run) puts the run on the list
      long_runs[atomicAdd(stats + 1, 1u)] = base + j;
  } else if (!skip) {
    const u32 p = (u32)l + rk;
    if (p != j) {
      s_src[p] = (unsigned short)j;
      ++moved;
    }
  }
}

// tags of window elements tid + k * FIN_THREADS, k in [K0, K0 + CNT), of an interior tile
template<int LOW_BITS, int K0, int CNT>
__device__ __forceinline__ void
finish_tags(const u64* __restrict__ kp, u32 tid, u32* s_tag)
{
  constexpr u32 LOW_MASK = (1u << LOW_BITS) - 1;
  u64 key[CNT], pk[CNT];
#pragma unroll
  for (int c = 0; c < CNT; ++c) {
    const u32 j = tid + (K0 + c) * FIN_THREADS;
    key[c] = kp[j];
    pk[c] = kp[(int)j - 1];
  }
#pragma unroll
  for (int c = 0; c < CNT; ++c) {
    const u32 j = tid + (K0 + c) * FIN_THREADS;
    const bool head = ((key[c] ^ pk[c]) >> LOW_BITS) != 0;
    s_tag[1 + j] = ((u32)key[c] & LOW_MASK) | (head ? FIN_HEAD : 0u);
  }
}

template<int LOW_BITS>
__global__ void __launch_bounds__(FIN_THREADS, 6)
segment_finish_kernel(u64* __restrict__ keys, u32* __restrict__ ids, u32 n, u32* __restrict__ stats,
                      u32* __restrict__ long_runs)
{
  static_assert(LOW_BITS <= 24, "tags keep the low bits next to two flag bits");
  // [0] left sentinel, [1 + j] window element j, then sentinels up to the end of the last warp window
  __shared__ u32 s_tag[1 + FIN_WINDOW + 32];
  __shared__ unsigned short s_src[FIN_WINDOW];
  constexpr u32 LOW_MASK = (1u << LOW_BITS) - 1;
  const u32 tid = threadIdx.x;
  const u32 base = blockIdx.x * FIN_TILE;
  const u32 valid = (n - base) < FIN_WINDOW ? (n - base) : FIN_WINDOW; // window elements that exist

  // ---- tags -------------------------------------------------------------------------------------------------
  for (u32 i = tid; i < FIN_WINDOW / 2; i += FIN_THREADS)
    reinterpret_cast<u32*>(s_src)[i] = 0xffffffffu; // FIN_KEEP everywhere
  if (tid < 32)
    s_tag[1 + FIN_WINDOW + tid] = FIN_HEAD;
  if (valid == FIN_WINDOW && base != 0) {
    // interior tile (all but the first and the last two): no bounds tests; three chunks keep 6 + 6 + 5 pairs of
    // loads in flight per thread without spilling registers
    const u64* kp = keys + base;
    finish_tags<LOW_BITS, 0, 6>(kp, tid, s_tag);
    finish_tags<LOW_BITS, 6, 6>(kp, tid, s_tag);
    finish_tags<LOW_BITS, 12, FIN_ITEMS - 12>(kp, tid, s_tag);
    if (tid == 0) // left sentinel: the low bits of the element before the tile (inversion test of element 0)
      s_tag[0] = ((u32)kp[-1] & LOW_MASK) | FIN_HEAD | FIN_FOREIGN;
  } else {
#pragma unroll 1
    for (u32 j = tid; j < FIN_WINDOW; j += FIN_THREADS) {
      const u32 g = base + j;
      const u64 key = j < valid ? keys[g] : ~0ull; // bit 63 set: top bits no MortonIndex64 has
      const u64 pk = (g > 0 && j <= valid) ? keys[g - 1] : ~0ull;
      const bool head = ((key ^ pk) >> LOW_BITS) != 0 || g == 0 || j >= valid;
      s_tag[1 + j] = ((u32)key & LOW_MASK) | (head ? FIN_HEAD : 0u);
      if (j == 0)
        s_tag[0] = ((u32)pk & LOW_MASK) | FIN_HEAD | FIN_FOREIGN;
    }
  }
  __syncthreads();

  // ---- where every element belongs --------------------------------------------------------------------------
  // A warp looks at 32 consecutive elements at a time and advances by FIN_STRIDE = 24: a run that starts in the
  // first 24 lanes and ends inside the 32 is ranked with shuffles, in as many uniform steps as the longest such
  // run of the window has elements (no divergent scans; runs of real clouds have a handful of elements).  Runs
  // that reach beyond the window take the scalar scan over the tags (finish_scan_run).  Every element is seen by
  // one or two windows and handled by exactly one: lanes 0..7 of a window are lanes 24..31 of the one before.
  u32 steps = 0, moved = 0, long_elements = 0;
  bool unsorted_long = false;
  const u32 lane = tid & 31, warp = tid >> 5;
  const u32 le = 0xffffffffu >> (31 - lane); // lanes <= mine
  if constexpr (LOW_BITS < 24) {
    // dense clouds (48 top bits sorted by the passes): nearly every run has one element, which two tag reads show
#pragma unroll 1
    for (u32 j = tid; j < valid; j += FIN_THREADS) {
      const u32 tag = s_tag[1 + j];
      if (!(tag & s_tag[2 + j] & FIN_HEAD))
        finish_scan_run<LOW_BITS>(j, tag, s_tag, s_src, steps, moved, unsorted_long, long_elements, base, stats,
                                  long_runs);
    }
  } else {
#pragma unroll 1
    for (u32 w0 = warp * FIN_STRIDE; w0 < valid; w0 += (FIN_THREADS / 32) * FIN_STRIDE) {
      const u32 j = w0 + lane;
      const u32 tag = s_tag[1 + j]; // elements behind the last valid one carry FIN_HEAD
      const u32 lo = tag & LOW_MASK;
      const u32 heads = __ballot_sync(0xffffffffu, (tag & FIN_HEAD) != 0);
      const u32 hb = heads & le, ha = heads & ~le;
      const bool live = j < valid;
      const int l_lane = 31 - __clz(hb); // -1: the run starts before the window
      const int r_lane = ha ? __ffs(ha) - 1 : 32;
      // runs of this window that the shuffle path can rank
      const bool mine = live && hb != 0 && l_lane < FIN_STRIDE;
      const bool fast = mine && ha != 0;
      // elements nobody else scans: runs that leave the window at its end, or entered it more than 8 lanes ago (the
      // previous window could not see their end either); in the first window every run that comes from the left
      const bool slow = live && ((mine && ha == 0) || (hb == 0 && (lane >= 32 - FIN_STRIDE || w0 == 0)));
      const int len = fast ? r_lane - l_lane : 1;
      const int maxlen = __reduce_max_sync(0xffffffffu, len);
      // (low bits, lane) as one number: an element precedes another iff its number is smaller (ties: earlier lane)
      const u32 val = (lo << 5) | lane;
      u32 rank = 0;
      for (int d = 1; d < maxlen; ++d) {
        int partner = (int)lane + d;
        if (partner >= r_lane)
          partner -= len;
        const u32 v = __shfl_sync(0xffffffffu, val, d < len ? partner : (int)lane);
        rank += v < val ? 1u : 0u;
      }
      if (fast) {
        steps += (u32)len - 1u;
        const u32 l = w0 + (u32)l_lane;
        const u32 p = l + rank;
        if (l < FIN_TILE && p != j) { // runs that start behind this tile belong to the next one
          s_src[p] = (unsigned short)j;
          ++moved;
        }
      }
      if (slow)
        finish_scan_run<LOW_BITS>(j, tag, s_tag, s_src, steps, moved, unsorted_long, long_elements, base, stats,
                                  long_runs);
    }
  }
  __syncthreads();

  // ---- pull ----------------------------------------------------------------------------------------------------
  u32 id[FIN_ITEMS];
#pragma unroll
  for (int k = 0; k < FIN_ITEMS; ++k) {
    const u32 src = s_src[tid + k * FIN_THREADS];
    id[k] = src != FIN_KEEP ? ids[base + src] : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < FIN_ITEMS; ++k) {
    const u32 j = tid + k * FIN_THREADS;
    const u32 src = s_src[j];
    if (src != FIN_KEEP) {
      const u32 g = base + j;
      keys[g] = (keys[g] & ~(u64)LOW_MASK) | (u64)(s_tag[1 + src] & LOW_MASK);
      ids[g] = id[k];
    }
  }
  if (unsorted_long)
    atomicOr(stats, FIN_UNSORTED);
  if (__any_sync(0xffffffffu, long_elements != 0)) {
    long_elements = __reduce_add_sync(0xffffffffu, long_elements);
    if (lane == 0)
      atomicAdd(stats, long_elements);
  }
  // work counters (feedback for the choice of first_pass): scan steps and moved elements
  steps = __reduce_add_sync(0xffffffffu, steps);
  moved = __reduce_add_sync(0xffffffffu, moved);
  if (lane == 0) {
    if (steps)
      atomicAdd(reinterpret_cast<unsigned long long*>(stats + 2), (unsigned long long)steps);
    if (moved)
      atomicAdd(reinterpret_cast<unsigned long long*>(stats + 4), (unsigned long long)moved);
  }
}

// Runs of more than FIN_LIMIT elements (thousands of returns from one pole, points piled up on the bounds by the
// clamp of index_point): one block per run on the list the finish kernel wrote.  The block finds the end of the run,
// returns if the run is in order already, and otherwise sorts it by its low bits with a stable LSD counting sort of
// its own: histogram by all threads, ranks by one warp walking the run in order (match_any groups equal digits),
// ping-pong with the same index range of the sort's second buffer pair.  Slow per element, but such runs hold a
// fraction of a per cent of a cloud; the caller takes the eight LSD passes instead when they hold more than 1/8.
template<int LOW_BITS>
__global__ void __launch_bounds__(256)
long_run_sort_kernel(u64* __restrict__ k0, u32* __restrict__ v0, u64* __restrict__ k1, u32* __restrict__ v1, u32 n,
                     const u32* __restrict__ long_runs)
{
  __shared__ u32 s_end;
  __shared__ u32 s_bin[256];
  const u32 tid = threadIdx.x, lane = tid & 31;
  const u32 start = long_runs[blockIdx.x];
  const u64 hi0 = k0[start] >> LOW_BITS;
  constexpr u32 LOW_MASK = (1u << LOW_BITS) - 1;
  if (tid == 0)
    s_end = n;
  __syncthreads();
  for (u32 i0 = start; i0 < n; i0 += 256) {
    const u32 i = i0 + tid;
    const bool diff = i < n && (k0[i] >> LOW_BITS) != hi0;
    if (diff)
      atomicMin(&s_end, i);
    if (__syncthreads_or(diff))
      break;
  }
  __syncthreads();
  const u32 end = s_end, len = end - start;
  bool descent = false;
  for (u32 i = start + 1 + tid; i < end; i += 256)
    descent |= ((u32)k0[i - 1] & LOW_MASK) > ((u32)k0[i] & LOW_MASK);
  if (!__syncthreads_or(descent))
    return;
  u64* ks = k0 + start;
  u32* vs = v0 + start;
  u64* kd = k1 + start;
  u32* vd = v1 + start;
  for (int shift = 0; shift < LOW_BITS; shift += 8) {
    s_bin[tid] = 0;
    __syncthreads();
    for (u32 i = tid; i < len; i += 256)
      atomicAdd(&s_bin[(u32)(ks[i] >> shift) & 255u], 1u);
    __syncthreads();
    if (tid < 32) {
      // exclusive scan of the 256 bins: 8 per lane
      u32 c[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        c[q] = s_bin[lane * 8 + q];
        sum += c[q];
      }
      u32 incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o)
          incl += up;
      }
      u32 run = incl - sum;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        s_bin[lane * 8 + q] = run;
        run += c[q];
      }
      __syncwarp();
      // stable scatter, 32 elements at a time in run order
      const u32 lt = lanemask_lt();
      for (u32 g0 = 0; g0 < len; g0 += 32) {
        const u32 i = g0 + lane;
        const bool valid = i < len;
        const u64 key = valid ? ks[i] : 0ull;
        const u32 val = valid ? vs[i] : 0u;
        const u32 d = valid ? ((u32)(key >> shift) & 255u) : (256u + lane);
        const u32 peers = __match_any_sync(0xffffffffu, d);
        const u32 before = __popc(peers & lt);
        u32 pos = 0;
        if (valid)
          pos = s_bin[d] + before;
        __syncwarp();
        if (valid && before == 0)
          s_bin[d] += __popc(peers);
        __syncwarp();
        if (valid) {
          kd[pos] = key;
          vd[pos] = val;
        }
      }
    }
    __syncthreads();
    u64* tk = ks;
    ks = kd;
    kd = tk;
    u32* tv = vs;
    vs = vd;
    vd = tv;
  }
  if (ks != k0 + start) // an odd number of passes ended in the second buffer pair
    for (u32 i = tid; i < len; i += 256) {
      k0[start + i] = ks[i];
      v0[start + i] = vs[i];
    }
}

// after the finish kernel flagged long unsorted runs: `n_runs` entries of `long_runs`
void
launch_long_run_sort(u64* keys0, u64* keys1, u32* vals0, u32* vals1, u64 n, int first_pass, const u32* long_runs,
                     u32 n_runs, cudaStream_t stream)
{
  if (n_runs == 0)
    return;
  switch (first_pass) {
    case 1: long_run_sort_kernel<8><<<n_runs, 256, 0, stream>>>(keys0, vals0, keys1, vals1, (u32)n, long_runs); break;
    case 2: long_run_sort_kernel<16><<<n_runs, 256, 0, stream>>>(keys0, vals0, keys1, vals1, (u32)n, long_runs); break;
    default: long_run_sort_kernel<24><<<n_runs, 256, 0, stream>>>(keys0, vals0, keys1, vals1, (u32)n, long_runs); break;
  }
}

// Run lengths of the SORTED keys, for the choice of the sort mode of the next batch (tiler.cu): counter
// [g * 8 + q] = number of elements i whose key equals key[i - 2^q] above bit 24 (g = 0) or bit 16 (g = 1), i.e. the
// elements that have at least 2^q predecessors in their run; counter [16] = elements looked at (a sample: every
// 8th chunk of 4 096 consecutive keys).  sum_q 2^max(q-1,0) * counter[q] bounds the comparisons
// the finish kernel would need; counter[7] says how many points sit in runs it would leave to long_run_sort_kernel.
#define RUN_STATS_CHUNK 4096 /* consecutive keys a block looks at */
#define RUN_STATS_EVERY 8    /* ... of every 8th chunk: the probe reads 1/8 of the keys (0.1 ms per 100 M) */
__global__ void __launch_bounds__(256)
run_stats_kernel(const u64* __restrict__ keys, u32 n, unsigned long long* __restrict__ out)
{
  __shared__ u32 s_cnt[17];
  if (threadIdx.x < 17)
    s_cnt[threadIdx.x] = 0;
  __syncthreads();
  u32 c[16];
#pragma unroll
  for (int q = 0; q < 16; ++q)
    c[q] = 0;
  const u64 first = (u64)blockIdx.x * RUN_STATS_EVERY * RUN_STATS_CHUNK;
  u32 seen = 0;
  for (u32 o = threadIdx.x; o < RUN_STATS_CHUNK; o += 256) {
    const u64 i = first + o;
    if (i >= n)
      break;
    ++seen;
    const u64 k = keys[i];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const u32 d = 1u << q;
      if (i >= d) {
        const u64 x = k ^ keys[i - d];
        c[q] += (x >> 24) == 0 ? 1u : 0u;
        c[8 + q] += (x >> 16) == 0 ? 1u : 0u;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const u32 v = __reduce_add_sync(0xffffffffu, c[q]);
    if ((threadIdx.x & 31) == 0 && v)
      atomicAdd(&s_cnt[q], v);
  }
  seen = __reduce_add_sync(0xffffffffu, seen);
  if ((threadIdx.x & 31) == 0)
    atomicAdd(&s_cnt[16], seen);
  __syncthreads();
  if (threadIdx.x < 17 && s_cnt[threadIdx.x])
    atomicAdd(&out[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// out17: 16 counters + the number of keys looked at
void
launch_run_stats(const u64* sorted_keys, u64 n, unsigned long long* out17, cudaStream_t stream)
{
  cudaMemsetAsync(out17, 0, 17 * sizeof(unsigned long long), stream);
  if (n == 0)
    return;
  const u64 span = (u64)RUN_STATS_EVERY * RUN_STATS_CHUNK;
  run_stats_kernel<<<(u32)((n + span - 1) / span), 256, 0, stream>>>(sorted_keys, (u32)n, out17);
}

void
launch_radix_sort_top(u64* keys0, u64* keys1, u32* vals0, u32* vals1, u64 n, int first_pass, u32* hist, u32* status,
                      u32* ticket, u32* stats, cudaStream_t stream, cudaEvent_t before_finish)
{
  if (n == 0)
    return;
  digit_base_kernel<<<1, RS_PASSES * 32, 0, stream>>>(hist);
  const bool from1 = sort_input_buffer_top(first_pass) == 1;
  run_passes(from1 ? keys1 : keys0, from1 ? vals1 : vals0, from1 ? keys0 : keys1, from1 ? vals0 : vals1, (u32)n,
             first_pass, true, hist, status, ticket, stream);
  if (before_finish)
    cudaEventRecord(before_finish, stream);
  if (first_pass == 0)
    return;
  const u32 tiles = (u32)((n + FIN_TILE - 1) / FIN_TILE);
  switch (first_pass) {
    case 1: segment_finish_kernel<8><<<tiles, FIN_THREADS, 0, stream>>>(keys0, vals0, (u32)n, stats, status); break;
    case 2: segment_finish_kernel<16><<<tiles, FIN_THREADS, 0, stream>>>(keys0, vals0, (u32)n, stats, status); break;
    default: segment_finish_kernel<24><<<tiles, FIN_THREADS, 0, stream>>>(keys0, vals0, (u32)n, stats, status); break;
  }
}

void
launch_radix_sort_again(u64* keys0, u64* keys1, u32* vals0, u32* vals1, u64 n, const u32* scanned_hist, u32* status,
                        u32* ticket, cudaStream_t stream)
{
  if (n == 0)
    return;
  const bool from1 = sort_input_buffer() == 1;
  if (from1) { // odd pass count (9-bit digits): the passes must start in buffer 1
    cudaMemcpyAsync(keys1, keys0, n * 8, cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(vals1, vals0, n * 4, cudaMemcpyDeviceToDevice, stream);
  }
  run_passes(from1 ? keys1 : keys0, from1 ? vals1 : vals0, from1 ? keys0 : keys1, from1 ? vals0 : vals1, (u32)n, 0,
             false, scanned_hist, status, ticket, stream);
}

// =============================================================================================
// K4  gathers
// =============================================================================================
// Random 24-byte records: ncu shows 145 B of DRAM reads per record (one 128-byte line per access, 1/8 of the
// records straddle two).  Neither cudaLimitMaxL2FetchGranularity = 32 nor ld.global.nc.L2::64B changes that on
// B200 (measured, round 2: 4.15 ms per 100 M records either way), so the loads stay plain.  More records in flight per
// thread do not help either (2 / 4 / 8 records per iteration: 4.21 / 5.00 / 4.57 ms instead of 4.10): with 17 GB of
// DRAM traffic in 4.1 ms the kernel sits at the rate the memory delivers random lines, not at a latency limit.
__global__ void __launch_bounds__(256)
gather_positions_kernel(const double* __restrict__ src, const u32* __restrict__ perm, u64 n, double* __restrict__ dst)
{
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
    const u64 s = perm[i];
    const double x = src[3 * s], y = src[3 * s + 1], z = src[3 * s + 2];
    dst[3 * i] = x;
    dst[3 * i + 1] = y;
    dst[3 * i + 2] = z;
  }
}

template<int W>
__global__ void __launch_bounds__(256)
gather_bytes_kernel(const unsigned char* __restrict__ src, const u32* __restrict__ perm, u64 n,
                    unsigned char* __restrict__ dst)
{
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
    const u64 s = perm[i];
#pragma unroll
    for (int k = 0; k < W; ++k)
      dst[i * W + k] = src[s * W + k];
  }
}

template<typename T>
__global__ void __launch_bounds__(256)
gather_words_kernel(const T* __restrict__ src, const u32* __restrict__ perm, u64 n, T* __restrict__ dst)
{
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256)
    dst[i] = src[perm[i]];
}

void
launch_gather_positions(const double* src, const u32* perm, u64 n, double* dst, cudaStream_t stream)
{
  if (n == 0)
    return;
  gather_positions_kernel<<<persistent_grid(n, 256, 8), 256, 0, stream>>>(src, perm, n, dst);
}

void
launch_gather_bytes(const void* src, const u32* perm, u64 n, u32 width, void* dst, cudaStream_t stream)
{
  if (n == 0)
    return;
  const int grid = persistent_grid(n, 256, 8);
  const unsigned char* s = static_cast<const unsigned char*>(src);
  unsigned char* d = static_cast<unsigned char*>(dst);
  switch (width) {
    case 1:
      gather_words_kernel<unsigned char><<<grid, 256, 0, stream>>>(s, perm, n, d);
      break;
    case 2:
      gather_words_kernel<unsigned short>
        <<<grid, 256, 0, stream>>>((const unsigned short*)src, perm, n, (unsigned short*)dst);
      break;
    case 3:
      gather_bytes_kernel<3><<<grid, 256, 0, stream>>>(s, perm, n, d);
      break;
    case 4:
      gather_words_kernel<u32><<<grid, 256, 0, stream>>>((const u32*)src, perm, n, (u32*)dst);
      break;
    case 8:
      gather_words_kernel<u64><<<grid, 256, 0, stream>>>((const u64*)src, perm, n, (u64*)dst);
      break;
    case 12:
      gather_bytes_kernel<12><<<grid, 256, 0, stream>>>(s, perm, n, d);
      break;
    case 16:
      gather_words_kernel<uint4><<<grid, 256, 0, stream>>>((const uint4*)src, perm, n, (uint4*)dst);
      break;
    default:
      break;
  }
}

__global__ void __launch_bounds__(256)
compose_ids_kernel(const u32* __restrict__ perm, const u32* __restrict__ idx, u64 n, u32* __restrict__ out)
{
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256)
    out[i] = perm[idx[i]];
}

void
launch_compose_ids(const u32* perm, const u32* idx, u64 n, u32* out, cudaStream_t stream)
{
  if (n == 0)
    return;
  compose_ids_kernel<<<persistent_grid(n, 256, 8), 256, 0, stream>>>(perm, idx, n, out);
}

// =============================================================================================
// FAST start-level support: boundaries of the 8^6 level-5 prefixes in the sorted keys
// =============================================================================================
__global__ void __launch_bounds__(256)
level5_bins_kernel(const u64* __restrict__ keys, u64 n, u32* __restrict__ bin_start)
{
  const u32 b = blockIdx.x * 256 + threadIdx.x;
  if (b > 262144u)
    return;
  const u64 target = (u64)b << 45; // 6 levels = 18 bits, 63 - 18 = 45
  u64 lo = 0, hi = n;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    if ((keys[mid] & SW_KEY_MASK) < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  bin_start[b] = (u32)lo;
}

void
launch_level5_bins(const u64* sorted_keys, u64 n, u32* bin_start, cudaStream_t stream)
{
  level5_bins_kernel<<<(262145 + 255) / 256, 256, 0, stream>>>(sorted_keys, n, bin_start);
}
