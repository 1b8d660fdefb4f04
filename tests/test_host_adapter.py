"""The C++ host side: swgpu_tiler.hpp (RAII layer over the C ABI) and TilingAlgorithmGPU.h (the
reference-side TilingAlgorithmBase adapter).  The adapter is compiled against the reference's own
headers (PointBuffer, AABB, Sampling, Range, ProgressReporter) plus tests/host_mock stand-ins for
the headers this image cannot build (taskflow etc.), linked with oracle/_ref/libswref.so for the
reference objects, and — on the GPU box — executed end to end and compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/schwarzwald"
DRIVER = os.path.join(ROOT, "oracle", "_ref", "adapter_driver")


def test_cxx_wrapper_compiles_standalone_and_fails_loudly_without_gpu(tmp_path):
    """swgpu_tiler.hpp needs nothing but the C ABI; without a GPU construction throws (no fallback)."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include "swgpu_tiler.hpp"
#include <cstdio>
#include <cstring>
int main() {
  if (swgpu::node_name(0755, 3) != "r755") return 3;
  double mn[3], mx[3]; const double rmin[3] = {0, 0, 0}, rmax[3] = {4, 4, 4};
  swgpu::node_bounds(5, 1, rmin, rmax, mn, mx);   // octant 5 = x and z upper halves
  if (mn[0] != 2 || mn[1] != 0 || mn[2] != 2 || mx[0] != 4 || mx[1] != 2 || mx[2] != 4) return 4;
  if (swgpu::sampling_from_name("JITTERED") != SW_JITTERED || swgpu::tiling_from_name("FAST") != SW_FAST) return 5;
  try { swgpu::sampling_from_name("NOPE"); return 6; } catch (const swgpu::Error&) {}
  try {
    swgpu::Tiler t(SW_RANDOM_GRID, SW_FAST, 0.01f, 100, 20000, rmin, rmax, 8);
    std::puts("created");
  } catch (const swgpu::Error& e) { std::printf("error %d\n", e.code); }
  return 0;
}''')
    exe = tmp_path / "t"
    lib = os.path.join(ROOT, "schwarzwald_b200")
    subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(lib, "host"), str(src), "-o", str(exe), "-L" + lib,
                    "-lswgpu", "-Wl,-rpath," + lib], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    import torch
    assert out.stdout.strip() == ("created" if torch.cuda.is_available() else "error 2")


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference headers")
def test_adapter_compiles_against_reference_headers(tmp_path):
    src = tmp_path / "a.cpp"
    src.write_text('#include "TilingAlgorithmGPU.h"\nint main() { return 0; }\n')
    inc = ["-I" + os.path.join(ROOT, "tests", "host_mock"), "-I" + os.path.join(ROOT, "oracle", "shim"),
           "-I" + REF + "/core", "-I" + REF + "/util", "-I/root/reference/lib/tl_expected",
           "-I/root/reference/lib/rapidjson/include", "-I" + os.path.join(ROOT, "schwarzwald_b200", "host")]
    out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w"] + inc + [str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def _driver_points(n, seed):
    """The xorshift64* generator of tests/host_mock/adapter_driver.cpp."""
    mask = (1 << 64) - 1
    state = (seed * 2654435761 + 88172645463325252) & mask
    out = np.empty((n, 3), np.float64)
    for i in range(n):
        for a in range(3):
            state ^= state >> 12
            state = (state ^ (state << 25)) & mask
            state ^= state >> 27
            out[i, a] = float(((state * 2685821657736338717) & mask) >> 44) * 100.0 / 1048576.0
    return out


def _fnv1a(buf):
    h = 1469598103934665603
    for b in buf:
        h = ((h ^ b) * 1099511628211) & ((1 << 64) - 1)
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("sampling,tiling", [("RANDOM_GRID", "FAST"), ("GRID_CENTER", "ACCURATE"),
                                             ("JITTERED", "FAST"), ("MIN_DISTANCE", "ACCURATE")])
def test_adapter_end_to_end_matches_oracle(port_oracle, sampling, tiling):
    """TilingAlgorithmGPU driven like Tiler::run drives V1/V3: node names, counts and the stored
    positions (through PointReference) equal the oracle's, progress counts every point once."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/adapter_driver was not built (needs /root/reference at build time)")
    from oracle import sworacle
    import schwarzwald_b200 as sw
    n, seed, max_pts, threads = 40_000, 5, 300, 2
    out = subprocess.run([DRIVER, str(n), str(seed), sampling, tiling, str(max_pts), str(threads)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "PROGRESS %d of %d" % (n, n)
    got = {}
    for ln in lines[:-1]:
        name, count, digest = ln.split()
        got[name] = (int(count), int(digest, 16))
    xyz = _driver_points(n, seed)
    bmin, bmax = np.zeros(3), np.full(3, 100.0)
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    params = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=max_pts, concurrency=threads)
    want = port_oracle.tile(params, xyz).as_dict()
    assert sorted(got) == sorted(want)
    for name, ids in want.items():
        assert got[name][0] == len(ids), name
        assert got[name][1] == _fnv1a(np.ascontiguousarray(xyz[ids.astype(np.int64)]).tobytes()), name


@pytest.mark.gpu
@pytest.mark.parametrize("sampling,tiling", [("RANDOM_GRID", "ACCURATE"), ("GRID_CENTER", "FAST"),
                                             ("JITTERED", "ACCURATE"), ("MIN_DISTANCE", "FAST")])
def test_adapter_multi_batch_matches_oracle(port_oracle, sampling, tiling):
    """internal_cache_size smaller than the cloud (the reference's default regime, executable/main.cpp:233-236):
    the adapter feeds every batch to the device-resident node store and hands the FINAL content of every node to
    the sink; it equals the oracle's tile_batches (tile_node with cached points, TilingAlgorithms.cpp:351-492)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/adapter_driver was not built (needs /root/reference at build time)")
    from oracle import sworacle
    import schwarzwald_b200 as sw
    n, seed, max_pts, threads, batch = 40_000, 7, 300, 2, 15_000
    out = subprocess.run([DRIVER, str(n), str(seed), sampling, tiling, str(max_pts), str(threads), str(batch)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "PROGRESS %d of %d" % (n, n)
    got = {}
    for ln in lines[:-1]:
        name, count, digest = ln.split()
        got[name] = (int(count), int(digest, 16))
    xyz = _driver_points(n, seed)
    bmin, bmax = np.zeros(3), np.full(3, 100.0)
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    params = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=max_pts, concurrency=threads)
    want = port_oracle.tile_batches(params, xyz, [15_000, 15_000, 10_000]).as_dict()
    assert sorted(got) == sorted(want)
    for name, ids in want.items():
        assert got[name][0] == len(ids), name
        assert got[name][1] == _fnv1a(np.ascontiguousarray(xyz[ids.astype(np.int64)]).tobytes()), name


@pytest.mark.gpu
@pytest.mark.parametrize("sampling,tiling", [("GRID_CENTER", "FAST"), ("JITTERED", "ACCURATE")])
def test_adapter_on_several_gpus_matches_oracle(port_oracle, sampling, tiling):
    """TilingAlgorithmGPU constructed with a device list: the single process shards the batch over 3 ranks
    (swgpu_multi_*; the GPUs of the box, shared when there are fewer) and hands the merged nodes to the sink."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/adapter_driver was not built (needs /root/reference at build time)")
    from oracle import sworacle
    import schwarzwald_b200 as sw
    n, seed, max_pts, threads = 60_000, 9, 300, 2
    env = dict(os.environ, SWGPU_TEST_GPUS=str(torch.cuda.device_count()))
    out = subprocess.run([DRIVER, str(n), str(seed), sampling, tiling, str(max_pts), str(threads), "0", "3"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "PROGRESS %d of %d" % (n, n)
    got = {}
    for ln in lines[:-1]:
        name, count, digest = ln.split()
        got[name] = (int(count), int(digest, 16))
    xyz = _driver_points(n, seed)
    bmin, bmax = np.zeros(3), np.full(3, 100.0)
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    params = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=max_pts, concurrency=threads)
    want = port_oracle.tile(params, xyz).as_dict()
    assert sorted(got) == sorted(want)
    for name, ids in want.items():
        assert got[name][0] == len(ids), name
        assert got[name][1] == _fnv1a(np.ascontiguousarray(xyz[ids.astype(np.int64)]).tobytes()), name
