// kernels_payload.cu — node payloads for the writer hand-off.
//
// After tiling, the reference's persist_points() walks every node's PointReferences one point at a
// time and converts the position for the output format.  Here the conversion runs over the whole
// node-major output at once, so that the host receives the bytes the writer stores:
//   PNTS  attributes::PositionAttribute::extractFromPoints   io/PNTSWriter.cpp:326-342
//         static_cast<float> of every coordinate (positions were shifted to the RTC centre before
//         indexing, process/TilerProcess.cpp:552-559)
//   LAS   LASPersistence::persist_points                     io/LASPersistence.h:119-131,160-163
//         header offset = node bounds min, scale = compute_las_scale_from_bounds(node bounds)
//         (io/LASPersistence.cpp:17-28), record X = laszip_set_coordinates' quantisation
//         I32_QUANTIZE((p - offset) / scale)  (LASzip, third-party, absent from the reference tree:
//         (n >= 0) ? (I32)(n + 0.5) : (I32)(n - 0.5))
// All HBM-bound: 4 B index + 4 B permutation + one 24-byte position gather in, 12 B out per point.
#include "swgpu_internal.cuh"

__device__ __forceinline__ const double*
payload_position(const double* __restrict__ xyz, const u32* __restrict__ perm, const u32* __restrict__ out_idx, u64 j)
{
  const u32 s = out_idx[j];
  const u64 id = perm ? perm[s] : s;
  return xyz + 3 * id;
}

__global__ void __launch_bounds__(256)
payload_pnts_kernel(const double* __restrict__ xyz, const u32* __restrict__ perm, const u32* __restrict__ out_idx,
                    u64 n_out, float* __restrict__ out)
{
  for (u64 j = (u64)blockIdx.x * 256 + threadIdx.x; j < n_out; j += (u64)gridDim.x * 256) {
    const double* p = payload_position(xyz, perm, out_idx, j);
    const double x = p[0], y = p[1], z = p[2];
    out[3 * j] = (float)x; // round to nearest even, as static_cast<float> does on the host
    out[3 * j + 1] = (float)y;
    out[3 * j + 2] = (float)z;
  }
}

__device__ __forceinline__ int
i32_quantize(double n)
{
  return (n >= 0) ? __double2int_rz(n + 0.5) : __double2int_rz(n - 0.5);
}

__global__ void __launch_bounds__(256)
payload_las_kernel(const double* __restrict__ xyz, const u32* __restrict__ perm, const u32* __restrict__ out_idx,
                   u64 n_out, const u64* __restrict__ node_first, u32 n_nodes, const double* __restrict__ node_hdr,
                   int* __restrict__ out)
{
  const u64 fmask = ~(1ull << 63);
  for (u64 j = (u64)blockIdx.x * 256 + threadIdx.x; j < n_out; j += (u64)gridDim.x * 256) {
    // node of output j: the last row whose first offset is <= j (empty rows share their successor's
    // offset and are skipped by taking the LAST such row)
    u32 lo = 0, hi = n_nodes;
    while (hi - lo > 1) {
      const u32 mid = (lo + hi) >> 1;
      if ((node_first[mid] & fmask) <= j)
        lo = mid;
      else
        hi = mid;
    }
    const double4 hdr = reinterpret_cast<const double4*>(node_hdr)[lo];
    const double* p = payload_position(xyz, perm, out_idx, j);
    const double x = p[0], y = p[1], z = p[2];
    out[3 * j] = i32_quantize((x - hdr.x) / hdr.w);
    out[3 * j + 1] = i32_quantize((y - hdr.y) / hdr.w);
    out[3 * j + 2] = i32_quantize((z - hdr.z) / hdr.w);
  }
}

static int
payload_grid(u64 n)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const u64 want = (n + 255) / 256;
  const u64 cap = (u64)sms * 8;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

void
launch_payload_pnts(const double* xyz, const u32* perm, const u32* out_idx, u64 n_out, float* out, cudaStream_t stream)
{
  if (n_out == 0)
    return;
  payload_pnts_kernel<<<payload_grid(n_out), 256, 0, stream>>>(xyz, perm, out_idx, n_out, out);
}

void
launch_payload_las(const double* xyz, const u32* perm, const u32* out_idx, u64 n_out, const u64* node_first,
                   u32 n_nodes, const double* node_hdr, int* out, cudaStream_t stream)
{
  if (n_out == 0 || n_nodes == 0)
    return;
  payload_las_kernel<<<payload_grid(n_out), 256, 0, stream>>>(xyz, perm, out_idx, n_out, node_first, n_nodes,
                                                              node_hdr, out);
}
