// Microbenchmark: SM throughput of the warp-level primitives a radix-sort ranking can be built from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_ops warp_ops.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32;
typedef unsigned long long u64;

__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ u32 match8_ballot(u32 d)
{
  u32 m = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const u32 v = __ballot_sync(0xffffffffu, bit);
    m &= bit ? v : ~v;
  }
  return m;
}

template<int MODE>
__global__ void __launch_bounds__(256) bench(u32* out, int iters, u32 seed)
{
  __shared__ u32 s_hist[8 * 256];
  for (int i = threadIdx.x; i < 2048; i += 256) s_hist[i] = 0;
  __syncthreads();
  u32 x = seed ^ (threadIdx.x * 2654435761u) ^ (blockIdx.x * 40503u);
  u32 acc = 0;
  u32* my = s_hist + (threadIdx.x >> 5) * 256;
  const u32 lt = lanemask_lt();
  for (int i = 0; i < iters; ++i) {
    x = x * 1664525u + 1013904223u;
    const u32 d = (x >> 13) & 255u;
    if (MODE == 0) acc += __match_any_sync(0xffffffffu, d);
    if (MODE == 1) acc += match8_ballot(d);
    if (MODE == 2) acc += __popc(x & lt);
    if (MODE == 3) acc += __shfl_sync(0xffffffffu, x, d & 31);
    if (MODE == 4) atomicAdd(&my[d], 1u);
    if (MODE == 5) { // v1 chain: match + LDS + STS
      const u32 peers = __match_any_sync(0xffffffffu, d);
      const u32 lower = __popc(peers & lt);
      const u32 base = my[d];
      __syncwarp();
      if (lower == 0) my[d] = base + __popc(peers);
      __syncwarp();
      acc += base + lower;
    }
    if (MODE == 6) { // ballot chain
      const u32 peers = match8_ballot(d);
      const u32 lower = __popc(peers & lt);
      const u32 base = my[d];
      __syncwarp();
      if (lower == 0) my[d] = base + __popc(peers);
      __syncwarp();
      acc += base + lower;
    }
    if (MODE == 7) acc += __ballot_sync(0xffffffffu, d & 1);
    if (MODE == 8) { acc += my[d]; }
    if (MODE == 9) { // atomicAdd with return by leader + shfl (v2 chain)
      const u32 peers = match8_ballot(d);
      const u32 lower = __popc(peers & lt);
      u32 base = 0;
      if (lower == 0) base = atomicAdd(&my[d], (u32)__popc(peers));
      base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
      acc += base + lower;
    }
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc + s_hist[threadIdx.x];
}

template<int MODE> void run(const char* name, u32* out)
{
  const int iters = 4096, blocks = 148 * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<blocks, 256>>>(out, 16, 1); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<MODE><<<blocks, 256>>>(out, iters, 7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_ops = (double)blocks * 8 * iters;
  // cycles per warp-op per SM at the nominal 1.965 GHz clock (8 CTAs x 8 warps resident per SM)
  const double cyc = ms * 1e-3 * 1.965e9 * 148 / warp_ops;
  printf("%-34s %8.3f ms  %7.2f SM-cycles per warp-op  (%.1f Gwarp-op/s)\n", name, ms, cyc, warp_ops / ms / 1e6);
}

int main()
{
  u32* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<2>("popc (baseline loop)", out);
  run<0>("match_any 8-bit random", out);
  run<1>("8 x ballot match", out);
  run<7>("single ballot", out);
  run<3>("shfl idx", out);
  run<4>("smem atomicAdd random digit", out);
  run<8>("LDS random digit", out);
  run<5>("v1 chain: match + LDS/STS", out);
  run<6>("ballot chain: 8 ballots + LDS/STS", out);
  run<9>("ballot + leader atomicAdd + shfl", out);
  return 0;
}
