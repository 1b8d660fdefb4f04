#!/usr/bin/env python
"""bench.py — points/sec of the tiler compute core (index + sort + LOD sample) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--points P]

A "step" is one pass of the hot path over one synthetic batch: Morton indexing, the radix sort,
the per-level sampling sweep and (FAST) the reconstruct of the skipped upper levels.  At N = 1 the
workload is BASELINE.json configs[1]: 100 M-point terrain-like cloud, RANDOM_GRID, FAST.  At N > 1
(torchrun, one rank per GPU) every rank generates 100 M points of an N x 100 M cloud, the points are
shuffled so that every GPU owns whole Morton-prefix subtrees (one kernel partitions the local points straight
into the peers' receive buffers over NVLink; NCCL carries the 16 KB histograms and the barriers), and each
GPU tiles its subtrees; `value` is all points / max-over-ranks device time ("weak" scaling).

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own CPU code path
(oracle/_ref when it was built from /root/reference, else the oracle port) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "points/sec tiled (index+sort+LOD sample)"
UNIT = "points/s"
WORKLOAD_POINTS = 100_000_000
SEED = 2
SAMPLING, TILING = "RANDOM_GRID", "FAST"
CONCURRENCY = 32  # num_indexing_threads of the reference run FAST's start-level estimate is matched to
CPU_SAMPLE_POINTS = 8_000_000


def workload_name(n_points):
    return ("synthetic %dM-point terrain-like cloud, RANDOM_GRID FAST, single batch"
            % (n_points // 1_000_000)) if n_points >= 1_000_000 else "synthetic %d-point terrain-like cloud" % n_points


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.

    NVML is polled from a thread every ~2 ms (a bench step is ~10 ms, `nvidia-smi -lms` cannot start
    that fast); `nvidia-smi --query-gpu` is the fallback when pynvml is unavailable."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bit masks (nvml.h)
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20,
                   "hw_thermal_slowdown": 0x40}

    def __init__(self, device_index=0):
        self.device_index = device_index
        self.samples = []  # (perf_counter, sm_mhz, reason bitmask)
        self.sm_max = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.mode = None

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.device_index])
            except (ValueError, IndexError):
                pass
        return self.device_index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
            self._sample_nvml()
        except Exception:
            self.mode = "smi"
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _sample_nvml(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        try:
            reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.samples.append((time.perf_counter(), sm, reasons))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self._nvml_index()), "--query-gpu=" + self.QUERY,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) < 9:
            return
        mask = 0
        for k, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if f[5 + k].lower().startswith("active"):
                mask |= self.REASON_BITS[name]
        self.samples.append((time.perf_counter(), float(f[1]), mask))
        self.sm_max = float(f[2])

    def _poll(self):
        while not self.stop_flag.is_set():
            try:
                if self.mode == "nvml":
                    self._sample_nvml()
                    time.sleep(0.002)
                else:
                    self._sample_smi()
            except Exception:
                time.sleep(0.05)

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in [t0, t1] (perf_counter times of the timed region)."""
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler not started"], "samples": 0}
        self.stop_flag.set()
        self.thread.join(timeout=15)
        inside = [x for x in self.samples if (t0 is None or x[0] >= t0) and (t1 is None or x[0] <= t1)]
        if not inside:
            inside = self.samples
        sm = sorted(x[1] for x in inside)
        mask = 0
        for x in inside:
            mask |= x[2]
        reasons = sorted(name for name, bit in self.REASON_BITS.items() if mask & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(sm), "source": self.mode}


def captured_traffic(n_points):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (same point count only)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if int(t["points"]) == int(n_points):
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        pass
    return None


def cpu_reference_run(n_points, steps, warmup, full_points, bounds=None):
    """Times the reference's CPU implementation of the path on a bounded sample (rank 0 only)."""
    import numpy as np
    import torch

    import schwarzwald_b200 as sw
    from oracle import sworacle
    from schwarzwald_b200 import synth

    kind = "reference" if sworacle.have_ref() else "port"
    orc = sworacle.Oracle("ref" if kind == "reference" else "port")
    # the reference's worker threads: indexing chunks and per-node tasks run on all host cores, its sort
    # is one std::sort (TilingAlgorithms.cpp:600-604,1289-1292)
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    orc.set_threads(cores)
    orc.set_reference_sort(True)  # std::sort, as the reference; the parity runs use the stable variant
    xyz = synth.generate("terrain", n_points, SEED, device="cpu").numpy()
    # the sample is tiled against the FULL cloud's bounds and spacing
    if bounds is None:  # generator extents: x,y span the full 10 km tile, z from the sample
        side = 10000.0
        tb_min = np.array([400000.0, 5600000.0, float(xyz[:, 2].min())])
        tb_max = np.array([400000.0 + side, 5600000.0 + side, float(xyz[:, 2].max())])
        bmin, bmax = sw.cubic_bounds(tb_min, tb_max)
    else:
        bmin, bmax = bounds
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    params = sworacle.make_params(SAMPLING, TILING, spacing, bmin, bmax, max_points_per_node=20000,
                                  concurrency=CONCURRENCY)
    times = []
    for it in range(warmup + steps):
        res = orc.tile(params, xyz)
        # the hot path only (index + sort + per-node sampling with the in-memory sink), timed inside the
        # library: building the PointBuffer from the numpy array and copying the results out are not part of it
        if it >= warmup:
            times.append(res.seconds)
    t = sum(times) / len(times)
    return {"value": n_points / t, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "first %d points of the same seeded terrain generator (of %d), same bounds/spacing, "
                      "single batch, in-memory sink; indexing and per-node tiling tasks on %d threads, one "
                      "std::sort as in the reference" % (n_points, full_points, cores),
            "seconds_per_pass": t, "nodes": int(len(res.nodes))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--points", type=int, default=WORKLOAD_POINTS, help="points per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 3))
        warm = min(args.warmup, 1)
        r = cpu_reference_run(CPU_SAMPLE_POINTS, steps, warm, args.points * max(1, args.gpus))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": r["seconds_per_pass"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/u64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.points), "sampling": SAMPLING, "tiling": TILING,
                           "concurrency": CONCURRENCY, "max_points_per_node": 20000},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch

    import schwarzwald_b200 as sw
    from schwarzwald_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the tiler kernels have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_local = args.points
    n_total = n_local * world

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic input, resident in HBM before the timed region ---------------------------------
    xyz = torch.empty((n_local, 3), dtype=torch.float64, device=dev)
    chunk = 1 << 24
    for s in range(0, n_local, chunk):
        m = min(chunk, n_local - s)
        xyz[s:s + m] = synth.terrain(m, seed=SEED, device=dev, start=rank * n_local + s)
    mn, mx = synth.tight_bounds(xyz)
    if world > 1:
        t = torch.tensor(np.concatenate([mn, -mx]), device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        t = t.cpu().numpy()
        mn, mx = t[:3], -t[3:]
    bmin, bmax = sw.cubic_bounds(mn, mx)
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)

    stream = torch.cuda.current_stream()
    if world > 1:
        from schwarzwald_b200.distributed import ShardedTiler
        tiler = ShardedTiler(SAMPLING, TILING, bmin, bmax, spacing, concurrency=CONCURRENCY, device=local_rank)
    else:
        tiler = sw.GpuTiler(SAMPLING, TILING, bmin, bmax, spacing, concurrency=CONCURRENCY, device=local_rank)
    tiler.set_stream(stream.cuda_stream)
    tiler.enable_timing(True)

    def step():
        tiler.build_execution_graph(xyz)
        tiler.finalize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sort_ms, launches = 0.0, 0
    stats = None
    barrier()
    t_begin = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        stats = tiler.stats()
        sort_ms += stats["ms_sort"]
        launches += stats["kernel_launches"]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    phase_ms = None
    if world > 1:  # one extra, untimed step with per-phase events (index / exchange / tile)
        tiler.profile = True
        step()
        phase_ms = {k: round(v, 3) for k, v in tiler.last["phase_ms"].items()}
        phase_ms["bytes_sent_off_gpu"] = tiler.last["bytes_sent_off_gpu"]
        phase_ms["exchange"] = tiler.last.get("exchange")
        tiler.profile = False
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, node table + point ids out, copies inside the timed region -----------
    e2e = None
    if world == 1 and not args.no_e2e:
        host = torch.empty((n_local, 3), dtype=torch.float64, pin_memory=True)
        host.copy_(xyz)
        host_np = host.numpy()
        nn, ni = tiler.result_size()
        ids_host = torch.empty(ni + 1024, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
        nodes_host = np.empty(nn + 1024, sw.tiler.NODE_DTYPE)
        e_steps = max(1, min(args.steps, 3))
        t_e2e, d2h = [], 0
        for it in range(1 + e_steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tiler.build_execution_graph(host_np)
            tiler.finalize()
            res = tiler.result(ids_out=ids_host, nodes_out=nodes_host)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it > 0:
                t_e2e.append(dt)
            d2h = int(res.ids.nbytes + res.nodes.nbytes)
        e2e = {"value": n_local / (sum(t_e2e) / len(t_e2e)), "unit": UNIT, "h2d_bytes_per_step": int(n_local * 24),
               "d2h_bytes_per_step": d2h, "steps": e_steps,
               "api": "swgpu_index_batch(host xyz) + swgpu_finalize + swgpu_get_nodes(host)"}
        del host
        # ---- the same call sequence fed with LAS record coordinates: 12 B/pt cross PCIe
        # instead of 24, the reader's int -> double conversion runs fused with the indexing kernel
        try:
            from schwarzwald_b200 import tiler as swt
            las_scale = np.array([0.001, 0.001, 0.001])
            las_offset = np.floor(mn)
            las_dev = torch.empty((n_local, 3), dtype=torch.int32, device=dev)
            off_t = torch.tensor(las_offset, device=dev)
            for s0 in range(0, n_local, chunk):
                m = min(chunk, n_local - s0)
                las_dev[s0:s0 + m] = torch.round((xyz[s0:s0 + m] - off_t) / 0.001).to(torch.int32)
            las_host = torch.empty((n_local, 3), dtype=torch.int32, pin_memory=True)
            las_host.copy_(las_dev)
            del las_dev
            las_np = las_host.numpy()
            tr = swt.las_transform(las_scale, las_offset, mn - 1.0, mx + 1.0)
            t_las = []
            for it in range(1 + e_steps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tiler.build_execution_graph_las(las_np, tr)
                tiler.finalize()
                res = tiler.result(ids_out=ids_host, nodes_out=nodes_host)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                if it > 0:
                    t_las.append(dt)
            e2e["las_input"] = {"value": n_local / (sum(t_las) / len(t_las)), "unit": UNIT,
                                "h2d_bytes_per_step": int(n_local * 12),
                                "d2h_bytes_per_step": int(res.ids.nbytes + res.nodes.nbytes), "steps": e_steps,
                                "api": "swgpu_index_batch_las(host int32 XYZ) + swgpu_finalize + swgpu_get_nodes(host)"}
            del las_host
        except Exception as ex:  # the primary e2e number above stands on its own
            e2e["las_input"] = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    # dominant kernel: one onesweep radix pass: 12 B read + 12 B written per point.  The pass count
    # (7 passes of 9 bits over the 63-bit keys) comes back through the library's byte accounting.
    n_passes = int(round((stats["bytes_sort"] / max(1, n_local) + 4) / 24.0))
    pass_ms = sort_ms / args.steps / n_passes
    achieved = (24.0 * n_local) / (pass_ms * 1e-3) / 1e9
    total_bytes = stats["bytes_index"] + stats["bytes_sort"] + stats["bytes_gather"] + stats["bytes_sample"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64/u64", "data": "synthetic",
        "config": {"workload": workload_name(n_local), "points_per_gpu": n_local, "sampling": SAMPLING, "tiling": TILING,
                   "concurrency": CONCURRENCY, "max_points_per_node": 20000, "seed": SEED,
                   "start_level": tiler.start_level(), "nodes": int(stats["n_nodes"]),
                   "output_ids": int(stats["n_output_ids"]),
                   "l2": "inputs (2.4 GB positions, 0.8 GB keys) are far larger than the 126 MB L2"},
        "roofline": {"bound": "hbm", "kernel": "onesweep_pass_kernel (%d launches per step)" % n_passes, "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": captured_traffic(n_local),
                     "algorithmic_bytes_per_launch": 24 * n_local,
                     "whole_step": {"algorithmic_bytes": int(total_bytes),
                                    "achieved": total_bytes / (ms_per_step * 1e-3) / 1e9,
                                    "frac": total_bytes / (ms_per_step * 1e-3) / 1e9 / peak}},
        "stage_ms": {k: stats[k] for k in ("ms_index", "ms_sort", "ms_gather", "ms_sample", "ms_total")},
        "gpu_launches": int(launches),
        "shuffle_phase_ms": phase_ms,
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(CPU_SAMPLE_POINTS, 1, 0, n_total, bounds=(bmin, bmax))
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
