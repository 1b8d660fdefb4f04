// Microbenchmark: what limits a single-pass decoupled-look-back kernel (the node_rle skeleton)?
// Variants switch off the ticket, the look-back, or both, and vary the tile size.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../schwarzwald_b200/csrc/common.cuh"

// look-back with a window of 32*K descriptors per round trip
template<int K>
__device__ __forceinline__ u64
lookback_wide(u64* status, u32 tile, u64 aggregate)
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
    u64 s[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const long long idx = pos - k * 32 - lane;
      s[k] = (idx >= 0) ? ld_relaxed_u64(status + idx) : SW_LB_PFX;
    }
    bool done = false;
#pragma unroll
    for (int k = 0; k < K; ++k) {
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
    if (done)
      break;
    pos -= 32 * K;
  }
  if (lane == 0)
    st_relaxed_u64(status + tile, SW_LB_PFX | ((prefix + aggregate) & SW_LB_MASK));
  return prefix;
}

template<int THREADS, int ITEMS, bool TICKET, bool LOOKBACK, bool VEC, int LBK = 0>
__global__ void __launch_bounds__(THREADS)
rle_kernel(const u64* __restrict__ keys, u64 count, int shift, u32* __restrict__ out, u32* __restrict__ n_out,
           u64* __restrict__ status, u32* __restrict__ ticket)
{
  constexpr int WARPS = THREADS / 32;
  constexpr int TILE = THREADS * ITEMS;
  __shared__ u32 s_slot;
  __shared__ u32 s_w[WARPS];
  __shared__ u64 s_prefix;
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u32 tile = blockIdx.x;
  if (TICKET) {
    if (threadIdx.x == 0)
      s_slot = atomicAdd(ticket, 1u);
    __syncthreads();
    tile = s_slot;
  }
  const u64 base = (u64)tile * TILE;
  u32 hmask[ITEMS];
  u32 wcount = 0;
  if (!VEC) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const u64 i = base + warp * (32 * ITEMS) + j * 32 + lane;
      bool head = false;
      if (i < count) {
        const u64 k = keys[i];
        head = (i == 0) || ((k >> shift) != (keys[i - 1] >> shift));
      }
      hmask[j] = __ballot_sync(0xffffffffu, head);
      wcount += __popc(hmask[j]);
    }
  } else {
    // one load per key; the previous key comes from the neighbouring lane
    u64 k[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const u64 i = base + warp * (32 * ITEMS) + j * 32 + lane;
      k[j] = (i < count) ? keys[i] : ~0ull;
    }
    const u64 wfirst = base + warp * (32 * ITEMS);
    u64 before = (lane == 0 && wfirst > 0 && wfirst < count) ? keys[wfirst - 1] : 0ull;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const u64 i = base + warp * (32 * ITEMS) + j * 32 + lane;
      u64 prev = __shfl_up_sync(0xffffffffu, k[j], 1);
      const u64 carry = (j == 0) ? before : __shfl_sync(0xffffffffu, k[j > 0 ? j - 1 : 0], 31);
      if (lane == 0)
        prev = carry;
      const bool head = (i < count) && ((i == 0) || ((k[j] >> shift) != (prev >> shift)));
      hmask[j] = __ballot_sync(0xffffffffu, head);
      wcount += __popc(hmask[j]);
    }
  }
  if (lane == 0)
    s_w[warp] = wcount;
  __syncthreads();
  u32 wexcl = 0, total = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    const u32 v = s_w[w];
    wexcl += (w < (int)warp) ? v : 0u;
    total += v;
  }
  u32 prefix = 0;
  if (LOOKBACK) {
    if (warp == 0) {
      const u64 p = LBK ? lookback_wide<LBK ? LBK : 1>(status, tile, (u64)total) : lookback_exclusive(status, tile, (u64)total);
      if (lane == 0)
        s_prefix = p;
    }
    __syncthreads();
    prefix = (u32)s_prefix;
  } else {
    prefix = tile * 3; // wrong on purpose: timing only
  }
  if (threadIdx.x == 0 && base + TILE >= count)
    *n_out = prefix + total;
  u32 run = prefix + wexcl;
  const u32 lt = lanemask_lt();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if ((hmask[j] >> lane) & 1u)
      out[(run + __popc(hmask[j] & lt)) & 0xFFFFF] = (u32)(base + warp * (32 * ITEMS) + j * 32 + lane);
    run += __popc(hmask[j]);
  }
}

__global__ void fill(u64* k, u64 n)
{
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    k[i] = (i * 11400714819323198485ull >> 40) + (i << 26); // increasing-ish: few heads
}

template<int THREADS, int ITEMS, bool TICKET, bool LOOKBACK, bool VEC, int LBK = 0>
void run(const char* name, const u64* keys, u64 n, u32* out, u32* n_out, u64* status, u32* ticket)
{
  constexpr int TILE = THREADS * ITEMS;
  const u32 tiles = (u32)((n + TILE - 1) / TILE);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemsetAsync(status, 0, (size_t)tiles * 8);
    cudaMemsetAsync(ticket, 0, 4);
    cudaEventRecord(e0);
    rle_kernel<THREADS, ITEMS, TICKET, LOOKBACK, VEC, LBK><<<tiles, THREADS>>>(keys, n, 45, out, n_out, status, ticket);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best)
      best = ms;
  }
  printf("%-44s tile=%5d  %7.3f ms  %7.1f GB/s  (%s)\n", name, TILE, best, n * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
  const u64 n = 100000000ull;
  u64 *keys, *status;
  u32 *out, *n_out, *ticket;
  cudaMalloc(&keys, n * 8);
  cudaMalloc(&status, (n / 1024 + 16) * 8);
  cudaMalloc(&out, (1 << 20) * 4 + 64);
  cudaMalloc(&n_out, 4);
  cudaMalloc(&ticket, 4);
  fill<<<1184, 256>>>(keys, n);
  cudaDeviceSynchronize();
  run<256, 8, true, true, false>("ticket + lookback (current)", keys, n, out, n_out, status, ticket);
  run<256, 8, false, true, false>("blockIdx + lookback", keys, n, out, n_out, status, ticket);
  run<256, 8, true, false, false>("ticket, no lookback", keys, n, out, n_out, status, ticket);
  run<256, 8, false, false, false>("blockIdx, no lookback", keys, n, out, n_out, status, ticket);
  run<256, 8, false, false, true>("blockIdx, no lookback, shfl prev", keys, n, out, n_out, status, ticket);
  run<256, 8, true, true, true>("ticket + lookback, shfl prev", keys, n, out, n_out, status, ticket);
  run<256, 16, true, true, true>("ticket + lookback, shfl prev", keys, n, out, n_out, status, ticket);
  run<512, 8, true, true, true>("ticket + lookback, shfl prev", keys, n, out, n_out, status, ticket);
  run<512, 16, true, true, true>("ticket + lookback, shfl prev", keys, n, out, n_out, status, ticket);
  run<1024, 8, true, true, true>("ticket + lookback, shfl prev", keys, n, out, n_out, status, ticket);
  run<128, 16, true, true, true>("ticket + lookback, shfl prev", keys, n, out, n_out, status, ticket);
  run<256, 16, false, false, true>("blockIdx, no lookback, shfl prev", keys, n, out, n_out, status, ticket);
  run<256, 8, true, true, true, 1>("wide lookback K=1", keys, n, out, n_out, status, ticket);
  run<256, 8, true, true, true, 2>("wide lookback K=2", keys, n, out, n_out, status, ticket);
  run<256, 8, true, true, true, 4>("wide lookback K=4", keys, n, out, n_out, status, ticket);
  run<256, 8, true, true, true, 8>("wide lookback K=8", keys, n, out, n_out, status, ticket);
  run<256, 16, true, true, true, 4>("wide lookback K=4", keys, n, out, n_out, status, ticket);
  run<256, 16, true, true, true, 8>("wide lookback K=8", keys, n, out, n_out, status, ticket);
  run<512, 8, true, true, true, 4>("wide lookback K=4", keys, n, out, n_out, status, ticket);
  return 0;
}
