// multi.cu — several GPUs from ONE host process behind the C ABI (swgpu_multi_*, include/swgpu.h).
//
// The reference is a single process: Tiler::build_execution_graph_for_indexing hands every batch to one
// TilingAlgorithmBase (core/process/Tiler.cpp:189-198, 499-527).  To let that process use all GPUs of a box, this
// driver runs the sharded pipeline of SURVEY.md section 8(e) with one host thread per GPU:
//
//   1. the batch is cut into n equal slices, slice r goes to GPU r (cudaMemcpy from the caller's PointBuffer)
//   2. local Morton keys (index_point clamps in place), coarse 4-level prefix histogram       [per GPU]
//   3. histograms meet on the host: splitters + the complete send/recv count matrix              [host, tiny]
//   4. partition AND exchange in one kernel: every point (+ its global id, + its attribute record) is written
//      straight into its destination GPU's receive buffer through peer access (cudaDeviceEnablePeerAccess,
//      NVLink)                                                                                 [per GPU]
//   5. every GPU tiles its whole Morton-prefix subtrees with the single-GPU pipeline; the two global quantities
//      (point counts of the nodes above the shard depth, level-5 counts for FAST's start level) and the
//      MIN_DISTANCE face exchange go through in-process collectives (pinned host staging + a barrier)
//   6. results: per-GPU node tables with global point ids, merged in rank (= Morton) order
//
// Everything on the data path is the same C ABI a torchrun rank uses (schwarzwald_b200/distributed.py); only the
// collectives differ (threads of one process instead of NCCL ranks).  `devices` may name a GPU more than once:
// the ranks then share that GPU (how the single-GPU test box exercises this file).
#include "swgpu_internal.cuh"

#include "../../include/swgpu.h"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// reusable barrier that can be aborted (a failing rank must not leave the others waiting)
class AbortableBarrier
{
public:
  explicit AbortableBarrier(unsigned n)
    : _n(n)
  {}
  bool wait()
  {
    std::unique_lock<std::mutex> lock(_m);
    if (_aborted)
      return false;
    const unsigned gen = _generation;
    if (++_arrived == _n) {
      _arrived = 0;
      ++_generation;
      _cv.notify_all();
      return true;
    }
    _cv.wait(lock, [&] { return _generation != gen || _aborted; });
    return !_aborted;
  }
  void abort()
  {
    std::lock_guard<std::mutex> lock(_m);
    _aborted = true;
    _cv.notify_all();
  }
  void reset()
  {
    std::lock_guard<std::mutex> lock(_m);
    _aborted = false;
    _arrived = 0;
  }

private:
  std::mutex _m;
  std::condition_variable _cv;
  unsigned _n, _arrived = 0, _generation = 0;
  bool _aborted = false;
};

struct DevMem
{
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap)
      return cudaSuccess;
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8;
    const cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess)
      cap = want;
    return e;
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct PinnedMem
{
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap)
      return cudaSuccess;
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const cudaError_t e = cudaMallocHost(&p, bytes);
    if (e == cudaSuccess)
      cap = bytes;
    return e;
  }
  void release()
  {
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

constexpr int COARSE_LEVELS = 4;
constexpr u32 COARSE_BINS = 4096;

} // namespace

struct swgpu_multi;

struct MultiRank
{
  swgpu_multi* owner = nullptr;
  u32 rank = 0;
  int device = 0;
  swgpu_handle h = nullptr;
  cudaStream_t stream = nullptr;
  DevMem xyz_in, attr_in, keys, bins, recv_xyz, recv_ids, recv_attr, face_recv;
  PinnedMem stage, sum, h_bins;
  u64 n_local = 0, n_shard = 0;
  int rc = SW_OK;
  std::string err;
  std::vector<sw_node> nodes;
  std::vector<u32> ids;
};

struct swgpu_multi
{
  sw_params prm{};
  u32 n = 0;
  std::vector<MultiRank> ranks;
  AbortableBarrier* barrier = nullptr;
  std::string err;
  u32 shard_levels = 0, max_shard_levels = 0;
  u32 first_prefix[SW_MAX_RANKS + 1] = {};
  std::vector<u64> count_matrix; // [source * n + destination]
  u64 n_global = 0;
  u32 attr_bytes = 0;
  bool batch_done = false;
  bool faces = true;
  // merged result
  std::vector<sw_node> nodes;
  std::vector<u32> ids;
  bool merged = false;
  u64 n_clamped = 0;
  int32_t start_level = -1;
  // shared pointers for the hooks
  void* stage_ptr[SW_MAX_RANKS] = {};
  u64 stage_bytes[SW_MAX_RANKS] = {};
};

namespace {

int
rank_fail(MultiRank& r, int code, const std::string& msg)
{
  r.rc = code;
  r.err = msg;
  r.owner->barrier->abort();
  return code;
}

#define RCK(expr)                                                                                                      \
  do {                                                                                                                 \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess) {                                                                                           \
      cudaGetLastError();                                                                                              \
      return rank_fail(r, _e == cudaErrorMemoryAllocation ? SW_ERR_OUT_OF_MEMORY : SW_ERR_CUDA,                        \
                       std::string("CUDA error: ") + cudaGetErrorString(_e) + " in " #expr);                           \
    }                                                                                                                  \
  } while (0)

#define LCK(expr)                                                                                                      \
  do {                                                                                                                 \
    const int _rc = (expr);                                                                                            \
    if (_rc != SW_OK)                                                                                                  \
      return rank_fail(r, _rc, swgpu_last_error(r.h));                                                                 \
  } while (0)

#define BARRIER()                                                                                                      \
  do {                                                                                                                 \
    if (!m->barrier->wait())                                                                                           \
      return SW_ERR_COLLECTIVE;                                                                                        \
  } while (0)

// swgpu_allreduce_u32_fn for the threads of one process: pinned staging, every rank sums all stages itself
int
allreduce_hook(void* ctx, uint32_t* counters_device, uint64_t count, void* cuda_stream)
{
  MultiRank& r = *static_cast<MultiRank*>(ctx);
  swgpu_multi* m = r.owner;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  cudaSetDevice(r.device);
  if (r.stage.ensure(count * 4) != cudaSuccess || r.sum.ensure(count * 4) != cudaSuccess)
    return 1;
  if (cudaMemcpyAsync(r.stage.p, counters_device, count * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    return 1;
  m->stage_ptr[r.rank] = r.stage.p;
  if (!m->barrier->wait())
    return 1;
  u32* sum = static_cast<u32*>(r.sum.p);
  std::memcpy(sum, m->stage_ptr[0], count * 4);
  for (u32 q = 1; q < m->n; ++q) {
    const u32* other = static_cast<const u32*>(m->stage_ptr[q]);
    for (u64 i = 0; i < count; ++i)
      sum[i] += other[i];
  }
  if (!m->barrier->wait()) // every rank has read every stage
    return 1;
  if (cudaMemcpyAsync(counters_device, sum, count * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    return 1;
  return 0;
}

// swgpu_allgatherv_fn for the threads of one process
int
allgatherv_hook(void* ctx, const void* send_device, uint64_t send_bytes, void** recv_device, uint64_t* recv_bytes,
                void* cuda_stream)
{
  MultiRank& r = *static_cast<MultiRank*>(ctx);
  swgpu_multi* m = r.owner;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  cudaSetDevice(r.device);
  if (r.stage.ensure(std::max<u64>(send_bytes, 8)) != cudaSuccess)
    return 1;
  if (send_bytes && (cudaMemcpyAsync(r.stage.p, send_device, send_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                     cudaStreamSynchronize(s) != cudaSuccess))
    return 1;
  m->stage_ptr[r.rank] = r.stage.p;
  m->stage_bytes[r.rank] = send_bytes;
  if (!m->barrier->wait())
    return 1;
  u64 total = 0;
  for (u32 q = 0; q < m->n; ++q)
    total += m->stage_bytes[q];
  if (r.face_recv.ensure(std::max<u64>(total, 8)) != cudaSuccess)
    return 1;
  u64 off = 0;
  for (u32 q = 0; q < m->n; ++q) {
    recv_bytes[q] = m->stage_bytes[q];
    if (m->stage_bytes[q] &&
        cudaMemcpyAsync(static_cast<char*>(r.face_recv.p) + off, m->stage_ptr[q], m->stage_bytes[q],
                        cudaMemcpyHostToDevice, s) != cudaSuccess)
      return 1;
    off += m->stage_bytes[q];
  }
  if (cudaStreamSynchronize(s) != cudaSuccess)
    return 1;
  *recv_device = r.face_recv.p;
  if (!m->barrier->wait()) // the stages may be overwritten by the next call
    return 1;
  return 0;
}

// one rank of one batch (runs on its own host thread)
int
rank_index_batch(MultiRank& r, double* xyz_host, u64 lo, u64 hi, const unsigned char* attr_host)
{
  swgpu_multi* m = r.owner;
  const u32 n_ranks = m->n;
  const u64 n = hi - lo;
  r.n_local = n;
  cudaSetDevice(r.device);
  cudaStream_t s = r.stream;
  // 1. this rank's slice of the batch
  RCK(r.xyz_in.ensure(std::max<u64>(n, 1) * 24));
  RCK(r.keys.ensure(std::max<u64>(n, 1) * 8));
  RCK(r.bins.ensure(COARSE_BINS * 4));
  RCK(r.h_bins.ensure(COARSE_BINS * 4));
  if (n)
    RCK(cudaMemcpyAsync(r.xyz_in.p, xyz_host + 3 * lo, n * 24, cudaMemcpyHostToDevice, s));
  if (m->attr_bytes) {
    RCK(r.attr_in.ensure(std::max<u64>(n, 1) * m->attr_bytes));
    if (n)
      RCK(cudaMemcpyAsync(r.attr_in.p, attr_host + lo * m->attr_bytes, n * m->attr_bytes, cudaMemcpyHostToDevice, s));
  }
  // 2. keys (index_point clamps in place) + coarse prefix histogram
  LCK(swgpu_morton_encode_device(r.h, static_cast<double*>(r.xyz_in.p), n, static_cast<uint64_t*>(r.keys.p)));
  uint64_t clamped = 0;
  LCK(swgpu_get_clamped_count(r.h, &clamped));
  if (clamped) { // index_point writes the clamped coordinates back into the PointBuffer
    RCK(cudaMemcpyAsync(xyz_host + 3 * lo, r.xyz_in.p, n * 24, cudaMemcpyDeviceToHost, s));
  }
  RCK(cudaMemsetAsync(r.bins.p, 0, COARSE_BINS * 4, s));
  LCK(swgpu_prefix_histogram_coarse_device(r.h, static_cast<const uint64_t*>(r.keys.p), n, COARSE_LEVELS,
                                           static_cast<uint32_t*>(r.bins.p)));
  RCK(cudaMemcpyAsync(r.h_bins.p, r.bins.p, COARSE_BINS * 4, cudaMemcpyDeviceToHost, s));
  RCK(cudaStreamSynchronize(s));
  m->stage_ptr[r.rank] = r.h_bins.p;
  m->stage_bytes[r.rank] = clamped;
  BARRIER();
  // 3. splitters and the count matrix (every rank computes the same numbers)
  std::vector<u64> coarse(COARSE_BINS, 0);
  for (u32 q = 0; q < n_ranks; ++q) {
    const u32* b = static_cast<const u32*>(m->stage_ptr[q]);
    for (u32 i = 0; i < COARSE_BINS; ++i)
      coarse[i] += b[i];
  }
  u64 n_global = 0;
  for (u64 c : coarse)
    n_global += c;
  const u32 unit = 1u << (3 * (6 - COARSE_LEVELS)); // level-5 prefixes per coarse bin
  const u32 shard_levels = std::min<u32>(m->max_shard_levels, COARSE_LEVELS);
  std::vector<u32> fine(SWGPU_PREFIX_BINS, 0);
  for (u32 i = 0; i < COARSE_BINS; ++i)
    fine[(size_t)i * unit] = (u32)coarse[i];
  u32 first_prefix[SW_MAX_RANKS + 1];
  if (swgpu_choose_splitters(fine.data(), n_ranks, shard_levels, first_prefix) != SW_OK)
    return rank_fail(r, SW_ERR_INVALID_ARGUMENT, "swgpu_choose_splitters failed");
  // count_matrix[q][d] = points of source q that go to destination d
  std::vector<u64> matrix((size_t)n_ranks * n_ranks, 0);
  for (u32 q = 0; q < n_ranks; ++q) {
    const u32* b = static_cast<const u32*>(m->stage_ptr[q]);
    for (u32 d = 0; d < n_ranks; ++d) {
      u64 c = 0;
      for (u32 i = first_prefix[d] / unit; i < first_prefix[d + 1] / unit; ++i)
        c += b[i];
      matrix[(size_t)q * n_ranks + d] = c;
    }
  }
  u64 id_base = 0, recv_total = 0, max_recv = 0;
  uint64_t dst_offsets[SW_MAX_RANKS] = {};
  for (u32 q = 0; q < r.rank; ++q)
    for (u32 d = 0; d < n_ranks; ++d)
      id_base += matrix[(size_t)q * n_ranks + d];
  for (u32 d = 0; d < n_ranks; ++d) {
    u64 tot = 0;
    for (u32 q = 0; q < n_ranks; ++q) {
      if (q < r.rank)
        dst_offsets[d] += matrix[(size_t)q * n_ranks + d];
      tot += matrix[(size_t)q * n_ranks + d];
    }
    if (d == r.rank)
      recv_total = tot;
    max_recv = std::max(max_recv, tot);
  }
  if (r.rank == 0) {
    m->n_global = n_global;
    m->shard_levels = shard_levels;
    std::memcpy(m->first_prefix, first_prefix, sizeof(first_prefix));
    m->count_matrix = matrix;
    u64 c = 0;
    for (u32 q = 0; q < n_ranks; ++q)
      c += m->stage_bytes[q];
    m->n_clamped = c;
  }
  r.n_shard = recv_total;
  // 4. receive buffers (peer-accessible), then partition + exchange in one kernel
  RCK(r.recv_xyz.ensure(std::max<u64>(recv_total, 1) * 24));
  RCK(r.recv_ids.ensure(std::max<u64>(recv_total, 1) * 4));
  if (m->attr_bytes)
    RCK(r.recv_attr.ensure(std::max<u64>(recv_total, 1) * m->attr_bytes));
  BARRIER(); // every rank has read the histograms and allocated its receive buffers
  void* peer_xyz[SW_MAX_RANKS];
  void* peer_ids[SW_MAX_RANKS];
  void* peer_attr[SW_MAX_RANKS];
  for (u32 q = 0; q < n_ranks; ++q) {
    peer_xyz[q] = m->ranks[q].recv_xyz.p;
    peer_ids[q] = m->ranks[q].recv_ids.p;
    peer_attr[q] = m->ranks[q].recv_attr.p;
  }
  LCK(swgpu_set_partition_attributes(r.h, m->attr_bytes ? r.attr_in.p : nullptr, m->attr_bytes, nullptr,
                                     m->attr_bytes ? peer_attr : nullptr));
  LCK(swgpu_partition_to_peers_device(r.h, static_cast<const uint64_t*>(r.keys.p), static_cast<const double*>(r.xyz_in.p),
                                      n, first_prefix, n_ranks, (u32)id_base, peer_xyz, peer_ids, dst_offsets, nullptr));
  LCK(swgpu_set_partition_attributes(r.h, nullptr, 0, nullptr, nullptr));
  RCK(cudaStreamSynchronize(s));
  BARRIER(); // every rank's points have arrived
  // 5. the single-GPU pipeline on the shard
  LCK(swgpu_set_shard(r.h, shard_levels, -1, allreduce_hook, &r, recv_total ? static_cast<const uint32_t*>(r.recv_ids.p) : nullptr));
  const bool md = m->prm.sampling == SW_MIN_DISTANCE || m->prm.sampling == SW_MIN_DISTANCE_FAST;
  if (md && m->faces)
    LCK(swgpu_set_shard_faces(r.h, first_prefix, n_ranks, r.rank, allgatherv_hook, &r));
  else
    LCK(swgpu_set_shard_faces(r.h, nullptr, 0, 0, nullptr, nullptr));
  LCK(swgpu_index_batch_device(r.h, recv_total ? static_cast<double*>(r.recv_xyz.p) : nullptr, recv_total));
  return SW_OK;
}

int
rank_finalize(MultiRank& r)
{
  LCK(swgpu_finalize(r.h));
  uint64_t nn = 0, ni = 0;
  LCK(swgpu_result_size(r.h, &nn, &ni));
  r.nodes.resize(nn);
  r.ids.resize(ni);
  LCK(swgpu_get_nodes(r.h, r.nodes.data(), r.ids.data()));
  return SW_OK;
}

template<typename Fn>
int
run_on_all_ranks(swgpu_multi* m, Fn fn)
{
  m->barrier->reset();
  std::vector<std::thread> threads;
  for (u32 q = 0; q < m->n; ++q) {
    m->ranks[q].rc = SW_OK;
    m->ranks[q].err.clear();
    threads.emplace_back([m, q, &fn]() {
      MultiRank& r = m->ranks[q];
      const int rc = fn(r);
      if (rc != SW_OK) {
        if (r.rc == SW_OK) { // released from a barrier by another rank's failure
          r.rc = rc;
          r.err = "aborted: another GPU failed";
        }
        m->barrier->abort();
      }
    });
  }
  for (auto& t : threads)
    t.join();
  // the first rank that failed for a reason of its own explains the error
  for (u32 q = 0; q < m->n; ++q)
    if (m->ranks[q].rc != SW_OK && m->ranks[q].rc != SW_ERR_COLLECTIVE) {
      m->err = "GPU rank " + std::to_string(q) + ": " + m->ranks[q].err;
      return m->ranks[q].rc;
    }
  for (u32 q = 0; q < m->n; ++q)
    if (m->ranks[q].rc != SW_OK) {
      m->err = "GPU rank " + std::to_string(q) + ": " + m->ranks[q].err;
      return m->ranks[q].rc;
    }
  return SW_OK;
}

// parts of one node concatenated in rank order = Morton order (nodes above the shard depth span GPUs)
void
merge_results(swgpu_multi* m)
{
  struct Part
  {
    u32 levels;
    u64 index;
    u32 rank;
    u32 row;
  };
  std::vector<Part> parts;
  for (u32 q = 0; q < m->n; ++q)
    for (u32 k = 0; k < m->ranks[q].nodes.size(); ++k)
      parts.push_back({ m->ranks[q].nodes[k].levels, m->ranks[q].nodes[k].index, q, k });
  std::stable_sort(parts.begin(), parts.end(), [](const Part& a, const Part& b) {
    if (a.levels != b.levels)
      return a.levels < b.levels;
    if (a.index != b.index)
      return a.index < b.index;
    return a.rank < b.rank;
  });
  m->nodes.clear();
  m->ids.clear();
  u64 total = 0;
  for (u32 q = 0; q < m->n; ++q)
    total += m->ranks[q].ids.size();
  m->ids.reserve(total);
  for (size_t i = 0; i < parts.size();) {
    size_t j = i;
    sw_node out{};
    out.levels = parts[i].levels;
    out.index = parts[i].index;
    out.first = m->ids.size();
    while (j < parts.size() && parts[j].levels == parts[i].levels && parts[j].index == parts[i].index) {
      const MultiRank& r = m->ranks[parts[j].rank];
      const sw_node& nd = r.nodes[parts[j].row];
      out.flags |= nd.flags;
      m->ids.insert(m->ids.end(), r.ids.begin() + nd.first, r.ids.begin() + nd.first + nd.count);
      ++j;
    }
    out.count = m->ids.size() - out.first;
    m->nodes.push_back(out);
    i = j;
  }
  m->merged = true;
}

} // namespace

extern "C" {

int
swgpu_multi_create(const sw_params* params, const int* devices, uint32_t n_devices, swgpu_multi_handle* out)
{
  if (!params || !devices || !out || n_devices == 0 || n_devices > SWGPU_MAX_RANKS)
    return SW_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  auto* m = new swgpu_multi();
  m->prm = *params;
  m->n = n_devices;
  m->ranks.resize(n_devices);
  m->barrier = new AbortableBarrier(n_devices);
  for (u32 q = 0; q < n_devices; ++q) {
    MultiRank& r = m->ranks[q];
    r.owner = m;
    r.rank = q;
    r.device = devices[q];
    const int rc = swgpu_create(params, devices[q], &r.h);
    if (rc != SW_OK) {
      swgpu_multi_destroy(m);
      return rc;
    }
    cudaSetDevice(r.device);
    if (cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking) != cudaSuccess) {
      swgpu_multi_destroy(m);
      return SW_ERR_CUDA;
    }
    swgpu_set_stream(r.h, r.stream);
  }
  // every GPU writes into every other GPU's receive buffers
  for (u32 a = 0; a < n_devices; ++a)
    for (u32 b = 0; b < n_devices; ++b) {
      if (devices[a] == devices[b])
        continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
      if (!can) {
        swgpu_multi_destroy(m);
        return SW_ERR_CUDA;
      }
      cudaSetDevice(devices[a]);
      const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        swgpu_multi_destroy(m);
        return SW_ERR_CUDA;
      }
      cudaGetLastError();
    }
  uint32_t msl = 0;
  swgpu_max_shard_levels(m->ranks[0].h, &msl);
  m->max_shard_levels = msl;
  *out = m;
  return SW_OK;
}

void
swgpu_multi_destroy(swgpu_multi_handle m)
{
  if (!m)
    return;
  for (MultiRank& r : m->ranks) {
    cudaSetDevice(r.device);
    if (r.stream)
      cudaStreamSynchronize(r.stream);
    if (r.h)
      swgpu_destroy(r.h);
    DevMem* bufs[] = { &r.xyz_in, &r.attr_in, &r.keys, &r.bins, &r.recv_xyz, &r.recv_ids, &r.recv_attr, &r.face_recv };
    for (DevMem* b : bufs)
      b->release();
    r.stage.release();
    r.sum.release();
    r.h_bins.release();
    if (r.stream)
      cudaStreamDestroy(r.stream);
  }
  delete m->barrier;
  delete m;
}

const char*
swgpu_multi_last_error(swgpu_multi_handle m)
{
  return m ? m->err.c_str() : "invalid handle";
}

int
swgpu_multi_set_min_distance_faces(swgpu_multi_handle m, int enable)
{
  if (!m)
    return SW_ERR_INVALID_ARGUMENT;
  m->faces = enable != 0;
  return SW_OK;
}

int
swgpu_multi_index_batch(swgpu_multi_handle m, double* xyz_host, uint64_t n, const void* attr_host, uint32_t attr_bytes)
{
  if (!m || (!xyz_host && n))
    return SW_ERR_INVALID_ARGUMENT;
  if (attr_host && (attr_bytes == 0 || attr_bytes > 16 || (attr_bytes & 3u))) {
    m->err = "attribute records must be 4, 8, 12 or 16 bytes per point";
    return SW_ERR_INVALID_ARGUMENT;
  }
  m->batch_done = false;
  m->merged = false;
  m->attr_bytes = attr_host ? attr_bytes : 0;
  if (n >= (1ull << 32)) {
    m->err = "global point ids are 32 bit: at most 2^32 - 1 points per batch";
    return SW_ERR_INVALID_ARGUMENT;
  }
  // reference behaviour on degenerate batches (TilingAlgorithms.cpp:253-259, threading/Parallel.h:181-186)
  if (n == 0) {
    m->err = "tile_internal_node: Got zero points to tile @ node r";
    return SW_ERR_EMPTY_NODE;
  }
  if (m->prm.tiling == SW_FAST && n < m->prm.concurrency) {
    m->err = "Can't scatter a range that has less than 'scatter_factor' elements!";
    return SW_ERR_TOO_FEW_POINTS;
  }
  if (m->max_shard_levels < 1) {
    m->err = "spacing too coarse to shard: a sampling cell would span GPUs";
    return SW_ERR_INVALID_ARGUMENT;
  }
  const unsigned char* attr = static_cast<const unsigned char*>(attr_host);
  const int rc = run_on_all_ranks(m, [&](MultiRank& r) {
    const u64 lo = n * r.rank / m->n, hi = n * (r.rank + 1) / m->n;
    return rank_index_batch(r, xyz_host, lo, hi, attr);
  });
  if (rc != SW_OK)
    return rc;
  swgpu_get_start_level(m->ranks[0].h, &m->start_level);
  m->batch_done = true;
  return SW_OK;
}

int
swgpu_multi_finalize(swgpu_multi_handle m)
{
  if (!m)
    return SW_ERR_INVALID_ARGUMENT;
  if (!m->batch_done)
    return SW_OK;
  const int rc = run_on_all_ranks(m, [&](MultiRank& r) { return rank_finalize(r); });
  if (rc != SW_OK)
    return rc;
  merge_results(m);
  return SW_OK;
}

int
swgpu_multi_result_size(swgpu_multi_handle m, uint64_t* n_nodes, uint64_t* n_point_ids)
{
  if (!m)
    return SW_ERR_INVALID_ARGUMENT;
  if (!m->merged) {
    m->err = "swgpu_multi_finalize has not run";
    return SW_ERR_STATE;
  }
  if (n_nodes)
    *n_nodes = m->nodes.size();
  if (n_point_ids)
    *n_point_ids = m->ids.size();
  return SW_OK;
}

int
swgpu_multi_get_nodes(swgpu_multi_handle m, sw_node* nodes, uint32_t* point_ids)
{
  if (!m)
    return SW_ERR_INVALID_ARGUMENT;
  if (!m->merged) {
    m->err = "swgpu_multi_finalize has not run";
    return SW_ERR_STATE;
  }
  if (nodes && !m->nodes.empty())
    std::memcpy(nodes, m->nodes.data(), m->nodes.size() * sizeof(sw_node));
  if (point_ids && !m->ids.empty())
    std::memcpy(point_ids, m->ids.data(), m->ids.size() * sizeof(u32));
  return SW_OK;
}

int
swgpu_multi_get_info(swgpu_multi_handle m, int32_t* start_level, uint32_t* shard_levels, uint64_t* n_clamped,
                     uint64_t* shard_points /* n_devices entries */)
{
  if (!m)
    return SW_ERR_INVALID_ARGUMENT;
  if (start_level)
    *start_level = m->start_level;
  if (shard_levels)
    *shard_levels = m->shard_levels;
  if (n_clamped)
    *n_clamped = m->n_clamped;
  if (shard_points)
    for (u32 q = 0; q < m->n; ++q)
      shard_points[q] = m->ranks[q].n_shard;
  return SW_OK;
}

// node-major attribute records of one GPU's part of the result (the attributes travelled with the points)
int
swgpu_multi_get_rank_attributes(swgpu_multi_handle m, uint32_t rank, sw_node* nodes, uint32_t* point_ids, void* attr_host)
{
  if (!m || rank >= m->n)
    return SW_ERR_INVALID_ARGUMENT;
  if (!m->merged || !m->attr_bytes) {
    m->err = "no finalized batch with attributes";
    return SW_ERR_STATE;
  }
  MultiRank& r = m->ranks[rank];
  cudaSetDevice(r.device);
  if (nodes && !r.nodes.empty())
    std::memcpy(nodes, r.nodes.data(), r.nodes.size() * sizeof(sw_node));
  if (point_ids && !r.ids.empty())
    std::memcpy(point_ids, r.ids.data(), r.ids.size() * sizeof(u32));
  if (attr_host && !r.ids.empty()) {
    DevMem tmp;
    if (tmp.ensure(r.ids.size() * (size_t)m->attr_bytes) != cudaSuccess) {
      m->err = "out of device memory";
      return SW_ERR_OUT_OF_MEMORY;
    }
    int rc = swgpu_gather_attribute_device(r.h, r.recv_attr.p, m->attr_bytes, tmp.p);
    if (rc == SW_OK && (cudaMemcpyAsync(attr_host, tmp.p, r.ids.size() * (size_t)m->attr_bytes, cudaMemcpyDeviceToHost,
                                        r.stream) != cudaSuccess ||
                        cudaStreamSynchronize(r.stream) != cudaSuccess))
      rc = SW_ERR_CUDA;
    tmp.release();
    if (rc != SW_OK) {
      m->err = swgpu_last_error(r.h);
      return rc;
    }
  }
  return SW_OK;
}

int
swgpu_multi_rank_result_size(swgpu_multi_handle m, uint32_t rank, uint64_t* n_nodes, uint64_t* n_point_ids)
{
  if (!m || rank >= m->n)
    return SW_ERR_INVALID_ARGUMENT;
  if (n_nodes)
    *n_nodes = m->ranks[rank].nodes.size();
  if (n_point_ids)
    *n_point_ids = m->ranks[rank].ids.size();
  return SW_OK;
}

} // extern "C"
