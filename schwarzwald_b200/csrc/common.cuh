// common.cuh — device-side building blocks shared by every kernel of libswgpu (sm_100a only).
//
//  * relaxed / volatile global accessors used by the decoupled look-back chains
//  * single-pass prefix "look-back" over tile descriptors (Merrill & Garland), used by the node
//    run-length encoder, the stable two-way compaction, the segmented arg-min and the onesweep
//    radix passes
//  * Morton / octree-bounds arithmetic shared by the indexing and sampling kernels
//
// Everything that mirrors reference arithmetic cites the reference file:line it must agree with
// bit for bit (paths relative to /root/reference/schwarzwald/).  The translation units including
// this header are compiled with -fmad=false: the reference is built without FMA contraction.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define SW_WARP 32

// ------------------------------------------------------------------------------------------
// memory-ordering helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u64
ld_relaxed_u64(const u64* p)
{
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void
st_relaxed_u64(u64* p, u64 v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u32
ld_relaxed_u32(const u32* p)
{
  u32 v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void
st_relaxed_u32(u32* p, u32 v)
{
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ u32
lanemask_lt()
{
  u32 m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// ------------------------------------------------------------------------------------------
// decoupled look-back over 64-bit tile descriptors
//   bits 63..62 : 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix
//   bits 61..0  : payload (callers may pack two 31-bit counters: they add without carrying into
//                 each other as long as each total stays below 2^31)
// Tiles are handed out through an atomic ticket, so every tile a CTA waits for belongs to a CTA
// that is already resident: the chain cannot dead-lock.
// ------------------------------------------------------------------------------------------
#define SW_LB_AGG (1ull << 62)
#define SW_LB_PFX (2ull << 62)
#define SW_LB_MASK ((1ull << 62) - 1)

// Called by all 32 lanes of ONE warp.  Publishes `aggregate` for `tile`, returns the exclusive
// prefix over tiles [0, tile) to every lane and publishes the inclusive prefix.
__device__ __forceinline__ u64
lookback_exclusive(u64* status, u32 tile, u64 aggregate)
{
  const u32 lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0)
      st_relaxed_u64(status, SW_LB_PFX | aggregate);
    return 0;
  }
  if (lane == 0)
    st_relaxed_u64(status + tile, SW_LB_AGG | aggregate);
  u64 prefix = 0;
  long long pos = (long long)tile - 1;
  while (true) {
    const long long idx = pos - lane;
    u64 s = SW_LB_PFX; // tiles before 0: inclusive prefix 0
    if (idx >= 0) {
      do {
        s = ld_relaxed_u64(status + idx);
      } while ((s >> 62) == 0);
    }
    const u32 pm = __ballot_sync(0xffffffffu, (s >> 62) == 2);
    const int first_p = pm ? (__ffs(pm) - 1) : 32;
    u64 v = ((int)lane <= first_p) ? (s & SW_LB_MASK) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
    prefix += v;
    if (pm)
      break;
    pos -= 32;
  }
  if (lane == 0)
    st_relaxed_u64(status + tile, SW_LB_PFX | ((prefix + aggregate) & SW_LB_MASK));
  return prefix;
}

// ------------------------------------------------------------------------------------------
// Morton arithmetic
// ------------------------------------------------------------------------------------------
// expand_bits_by_3(uint64_t), core/util/stuff.h:207-221 (21 input bits)
__host__ __device__ __forceinline__ u64
expand_bits_by_3(u64 v)
{
  v &= 0x1FFFFFull;
  v = (v | (v << 32)) & 0x001F00000000FFFFull;
  v = (v | (v << 16)) & 0x001F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

// contract_bits_by_3, core/util/stuff.h:223-234
__host__ __device__ __forceinline__ u64
contract_bits_by_3(u64 v)
{
  v &= 0x1249249249249249ull;
  v = (v | (v >> 2)) & 0x10C30C30C30C30C3ull;
  v = (v | (v >> 4)) & 0x100F00F00F00F00Full;
  v = (v | (v >> 8)) & 0x001F0000FF0000FFull;
  v = (v | (v >> 16)) & 0x001F00000000FFFFull;
  v = (v | (v >> 32)) & 0x00000000001FFFFFull;
  return v;
}

// the same for values below 2^30 (10 result bits) in 32-bit arithmetic: half the instructions on the GPU
__host__ __device__ __forceinline__ u32
contract_bits_by_3_u32(u32 v)
{
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030C30C3u;
  v = (v | (v >> 4)) & 0x0300F00Fu;
  v = (v | (v >> 8)) & 0x030000FFu;
  v = (v | (v >> 16)) & 0x000003FFu;
  return v;
}

// Dataset bounds + the per-axis scale of calculate_morton_index<21>
// (core/tiling/OctreeAlgorithms.h:69-72: scale = 2^21 / extent, computed once on the host in
// double precision exactly as the reference does per point).
struct SwBounds
{
  double min[3];
  double max[3];
  double scale[3];
};

// Key shift that keeps `levels` octree levels of a 63-bit key (levels = 0 -> everything equal).
__host__ __device__ __forceinline__ int
shift_for_levels(int levels)
{
  return 3 * (21 - levels);
}

// bit 63 of a working key is never part of a MortonIndex64 (63 bits)
#define SW_KEY_MASK 0x7FFFFFFFFFFFFFFFull

// calculate_morton_index<21> (tiling/OctreeAlgorithms.h:64-87) against any box `b` (min + scale = 2^21 / extent)
__device__ __forceinline__ u64
morton_from_position(double x, double y, double z, const SwBounds& b)
{
  // (p - min) * scale: subtract then multiply, two roundings (no FMA), then truncate toward zero
  // and cap at 2^21 - 1 (OctreeAlgorithms.h:71-79).
  const double nx = (x - b.min[0]) * b.scale[0];
  const double ny = (y - b.min[1]) * b.scale[1];
  const double nz = (z - b.min[2]) * b.scale[2];
  const u64 cap = (1u << 21) - 1;
  u64 bx = __double2ull_rz(nx);
  u64 by = __double2ull_rz(ny);
  u64 bz = __double2ull_rz(nz);
  bx = bx < cap ? bx : cap;
  by = by < cap ? by : cap;
  bz = bz < cap ? bz : cap;
  return expand_bits_by_3(bz) | (expand_bits_by_3(by) << 1) | (expand_bits_by_3(bx) << 2);
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

// the same recurrence continued from bounds that already hold `from_depth` levels
__device__ __forceinline__ void
bounds_continue(u64 key, int from_depth, int depth, double mn[3], double mx[3])
{
  for (int level = from_depth; level < depth; ++level) {
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
