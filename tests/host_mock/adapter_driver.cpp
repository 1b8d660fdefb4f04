// adapter_driver.cpp — runs schwarzwald_b200/host/TilingAlgorithmGPU.h the way Tiler::run does
// (core/process/Tiler.cpp:499-527, 284): reference PointBuffer + SamplingStrategy + AABB types,
// build_execution_graph on a taskflow, finalize, per-node persist_points into a sink.
// TEST INFRASTRUCTURE: compiled against /root/reference headers (+ tests/host_mock stand-ins) and
// linked with oracle/_ref/libswref.so for the reference's PointBuffer / Sampling objects.
// Prints one line per persisted node: "<name> <count> <fnv1a of the stored positions>".
#include "TilingAlgorithmGPU.h"

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <cstdlib>

static uint64_t
fnv1a(const void* data, size_t bytes, uint64_t h = 1469598103934665603ull)
{
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < bytes; ++i) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

int
main(int argc, char** argv)
{
  if (argc < 7) {
    std::fprintf(stderr,
                 "usage: adapter_driver n seed SAMPLING TILING max_points_per_node indexing_threads [batch_size] [n_gpu_ranks]\n");
    return 2;
  }
  const size_t n = std::strtoull(argv[1], nullptr, 10);
  uint64_t state = std::strtoull(argv[2], nullptr, 10) * 2654435761ull + 88172645463325252ull;
  const std::string sampling = argv[3], tiling = argv[4];
  const size_t max_points = std::strtoull(argv[5], nullptr, 10);
  const uint32_t threads = static_cast<uint32_t>(std::strtoul(argv[6], nullptr, 10));
  // internal_cache_size: Tiler::run hands the algorithm at most this many points per batch (Tiler.cpp:499-527)
  const size_t batch = argc > 7 ? std::strtoull(argv[7], nullptr, 10) : 0;
  // > 1: the adapter shards a single-batch run over that many ranks (GPUs 0, 1, ... modulo the GPUs of the box)
  const size_t gpu_ranks = argc > 8 ? std::strtoull(argv[8], nullptr, 10) : 1;

  std::vector<Vector3<double>> positions(n);
  for (auto& p : positions) { // xorshift64*, coordinates on a millimetre lattice in [0, 100) m
    double c[3];
    for (double& v : c) {
      state ^= state >> 12;
      state ^= state << 25;
      state ^= state >> 27;
      v = static_cast<double>((state * 2685821657736338717ull) >> 44) * 100.0 / 1048576.0;
    }
    p = Vector3<double>(c[0], c[1], c[2]);
  }
  PointBuffer buffer(n, std::move(positions));
  const AABB bounds{ { 0, 0, 0 }, { 100, 100, 100 } };

  TilerMetaParameters meta{};
  meta.spacing_at_root = static_cast<float>(bounds.extent().length() / 250.0); // TilerProcess.cpp:598-604
  meta.max_depth = 100;
  meta.max_points_per_node = max_points;
  meta.internal_cache_size = batch ? batch : n + 1;
  meta.tiling_strategy = tiling == "FAST" ? TilingStrategy::Fast : TilingStrategy::Accurate;

  // the switch of TilerProcess::make_sampling_strategy (core/process/TilerProcess.cpp:491-516)
  SamplingStrategy strategy = RandomSortedGridSampling{ max_points };
  if (sampling == "GRID_CENTER")
    strategy = GridCenterSampling{ max_points };
  else if (sampling == "MIN_DISTANCE")
    strategy = PoissonDiskSampling{ max_points };
  else if (sampling == "JITTERED")
    strategy = JitteredSampling{ max_points };
  ProgressReporter progress;
  progress.register_progress_counter<size_t>(progress::INDEXING, n);
  PointsPersistence sink;
  try {
    std::vector<int> devices;
    int n_gpus = 1;
    if (gpu_ranks > 1) {
      const char* env = std::getenv("SWGPU_TEST_GPUS");
      n_gpus = env ? std::max(1, std::atoi(env)) : 1;
    }
    for (size_t q = 0; q < gpu_ranks; ++q)
      devices.push_back(static_cast<int>(q % static_cast<size_t>(n_gpus)));
    TilingAlgorithmGPU algorithm(strategy, &progress, sink, meta, devices);
    const size_t step = batch ? batch : n;
    for (size_t lo = 0; lo < n; lo += step) { // Tiler::run: one execution graph per batch, run to completion
      // Tiler reuses its point caches: the batch lives in its own buffer that is gone after the batch
      const size_t hi = std::min(n, lo + step);
      std::vector<PointBuffer::PointReference> refs(std::begin(buffer) + lo, std::begin(buffer) + hi);
      PointBuffer cache{ gsl::span<PointBuffer::PointReference>{ refs.data(), refs.size() } };
      tf::Taskflow taskflow;
      algorithm.build_execution_graph({ std::begin(cache), std::end(cache) }, bounds, threads, taskflow);
      for (auto& work : taskflow.work) // the executor
        work();
    }
    algorithm.finalize(bounds);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 1;
  }
  for (const auto& [name, pts] : sink.nodes)
    std::printf("%s %zu %016" PRIx64 "\n", name.c_str(), pts.size(), fnv1a(pts.data(), pts.size() * sizeof(pts[0])));
  std::printf("PROGRESS %zu of %zu\n", progress.get_progress<size_t>(progress::INDEXING), n);
  return 0;
}
