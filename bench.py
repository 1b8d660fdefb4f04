#!/usr/bin/env python
"""bench.py — points/sec of the tiler compute core (index + sort + LOD sample) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1|c2|c3|c4|c5] [--impl reference]

A "step" is one pass of the hot path over one synthetic batch: Morton indexing, the radix sort, the per-level
sampling sweep, (FAST) the reconstruct of the skipped upper levels, and the hand-off — node table on the host,
node-major ORIGINAL point ids on the device.  The workload is a BASELINE.json config (schwarzwald_b200/workloads.py):

    N = 1 (default)   c2: 100 M-point terrain-like cloud, RANDOM_GRID FAST (the config the metric is quoted on)
    N > 1 (default)   c3: 500 M-point urban-like cloud, JITTERED FAST, STRONG-scaled: the same 500 M points on
                      N GPUs (torchrun, one rank per GPU; every rank generates its 1/N slice, the points are
                      shuffled over NVLink so that every GPU owns whole Morton-prefix subtrees)
    --config c4       1 B points, MIN_DISTANCE ACCURATE (8 GPUs);  c5: 2 B skewed points, GRID_CENTER FAST;  c1: 10 M

After the timed region every rank checks its result against the CPU oracle (untimed `parity` leg: the whole
result at c1/c2 on one GPU, a fixed sample of Morton-prefix subtrees otherwise).  Prints ONE JSON line (rank 0).
`--impl reference` times the reference's own CPU code path (oracle/_ref when it was built from /root/reference,
else the oracle port) on the same config, full size where host memory and a few minutes allow.
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
NVLINK_PEAK_GBS = 900.0  # NVLink 5, per direction per GPU (B200_PROFILING.md)
# reference arm: K + W passes over one PointBuffer; the sample is the full cloud unless the passes would not end
# within the time budget (rate from a calibration pass on this box) or host memory is too small
REF_TIME_BUDGET_S = 420.0   # all K + W passes of the reference arm together
REF_RATE_DERATE = 0.8       # a big cloud tiles slower per point than the calibration sample (n log n sort, caches)
REF_BYTES_PER_POINT = 130   # numpy input + PointBuffer copy + PointReference + IndexedPoint64 + sampling scratch


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


def same_config_n1(cfg_name):
    """points/s of the SAME config on one GPU, from the committed profile of this round (None if absent)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_%s_n1.json" % cfg_name)))
        return {"value": d["value"], "ms_per_step": d["ms_per_step"], "source": "profiles/r02_bench_%s_n1.json" % cfg_name}
    except Exception:
        return None


def config_block(cfg_name, cfg, n_total, extra=None):
    """The `config` object both arms print (same keys and values for the same workload)."""
    from schwarzwald_b200 import workloads
    c = {"workload": cfg["workload"] if n_total == cfg["points"] else
         cfg["workload"].replace("synthetic", "synthetic (%d points of the)" % n_total, 1),
         "name": cfg_name, "points": int(n_total), "sampling": cfg["sampling"], "tiling": cfg["tiling"],
         "concurrency": cfg["concurrency"], "max_points_per_node": workloads.MAX_POINTS_PER_NODE, "seed": cfg["seed"],
         "l2": "inputs (24 B position + 8 B key per point, >= 0.3 GB per GPU) are far larger than the 126 MB L2"}
    if extra:
        c.update(extra)
    return c


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 64 << 30


# --------------------------------------------------------------------------------------------------------------
# the cloud: generated on the device in chunks; bounds reduced over the FULL cloud
# --------------------------------------------------------------------------------------------------------------
def full_cloud_tight_bounds(cfg, n_total, device, start=0, count=None, keep=None):
    """min / max of points [start, start + count) of the config's cloud (chunked generation).  keep: optional
    (start, count) slice of the generated points to return alongside (a (count, 3) tensor)."""
    import torch
    from schwarzwald_b200 import workloads
    count = n_total if count is None else count
    mn = torch.full((3,), float("inf"), dtype=torch.float64, device=device)
    mx = torch.full((3,), float("-inf"), dtype=torch.float64, device=device)
    kept = None
    if keep is not None:
        kept = torch.empty((keep[1], 3), dtype=torch.float64, device=device)
    chunk = 1 << 24
    for s in range(start, start + count, chunk):
        m = min(chunk, start + count - s)
        pts = workloads.generate_slice(cfg, s, m, device, chunk)
        mn = torch.minimum(mn, pts.amin(dim=0))
        mx = torch.maximum(mx, pts.amax(dim=0))
        if keep is not None:
            lo, hi = max(s, keep[0]), min(s + m, keep[0] + keep[1])
            if lo < hi:
                kept[lo - keep[0]:hi - keep[0]] = pts[lo - s:hi - s]
        del pts
    return mn.cpu().numpy(), mx.cpu().numpy(), kept


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation on the host cores
# --------------------------------------------------------------------------------------------------------------
def make_oracle(cfg, spacing, bmin, bmax):
    """The reference's CPU implementation (oracle/_ref when it was built from /root/reference, else the port)
    configured like the reference run: all host cores, one std::sort."""
    from oracle import sworacle
    from schwarzwald_b200 import workloads
    kind = "reference" if sworacle.have_ref() else "port"
    orc = sworacle.Oracle("ref" if kind == "reference" else "port")
    # the reference's worker threads: indexing chunks and per-node tasks run on all host cores, its sort
    # is one std::sort (TilingAlgorithms.cpp:600-604,1289-1292)
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    orc.set_threads(cores)
    orc.set_reference_sort(True)  # std::sort, as the reference; parity runs use the stable variant
    params = sworacle.make_params(cfg["sampling"], cfg["tiling"], spacing, bmin, bmax,
                                  max_points_per_node=workloads.MAX_POINTS_PER_NODE, concurrency=cfg["concurrency"])
    return orc, params, kind, cores


def cpu_reference_run(cfg, xyz_np, bmin, bmax, spacing, steps, warmup, sample_note, keep_result=False):
    """Times the reference's CPU implementation of the path (rank 0 only) on `xyz_np`: `warmup` untimed and `steps`
    timed passes over one PointBuffer.  The time of a pass is taken inside the library around index + sort + per-node
    sampling with the in-memory sink; building the PointBuffer and copying results out are not part of it."""
    orc, params, kind, cores = make_oracle(cfg, spacing, bmin, bmax)
    res = orc.tile(params, xyz_np, passes=warmup + steps, copy=False, want_keys=False)
    times = [float(x) for x in res.pass_seconds[warmup:]]
    t = sum(times) / len(times)
    n = len(xyz_np)
    out = {"value": n / t, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": "%s; same bounds/spacing as the full cloud, single batch, in-memory sink; indexing and "
                     "per-node tiling tasks on %d threads, one std::sort as in the reference" % (sample_note, cores),
           "seconds_per_pass": t, "nodes": int(len(res.nodes)), "points": int(n), "duplicate_keys": int(res.duplicate_keys)}
    if keep_result:
        out["_result"] = res
        out["_oracle"] = orc
    return out


def reference_sample_size(cfg, n_total, passes, rate, budget_s=REF_TIME_BUDGET_S):
    """Points of one reference pass: the full cloud unless host memory or the time budget of the whole
    `passes`-pass run forbid it.  `rate`: points/s measured by a small calibration pass on this box."""
    by_mem = int(0.7 * mem_available_bytes() / REF_BYTES_PER_POINT)
    by_time = int(budget_s * rate * REF_RATE_DERATE / max(1, passes))
    n = min(n_total, by_mem, by_time)
    if n < n_total:
        n = max(1_000_000, n // 1_000_000 * 1_000_000)
    return min(n, n_total)


def run_reference_arm(args, cfg_name, cfg, n_total):
    import numpy as np
    import torch
    from schwarzwald_b200 import workloads

    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    steps, warm = max(1, args.steps), max(0, args.warmup)
    if dev.type == "cuda":  # input generation only: the measured code runs on the host cores
        mn, mx, _ = full_cloud_tight_bounds(cfg, n_total, dev)
    else:  # no GPU: bounds from the leading points
        head = workloads.generate_slice(cfg, 0, min(n_total, 4_000_000), dev)
        mn, mx = head.amin(dim=0).numpy(), head.amax(dim=0).numpy()
    bmin, bmax, spacing, centre = workloads.finish_bounds(cfg, mn, mx)

    def first_points(n):
        pts = workloads.apply_pre_transform(cfg, workloads.generate_slice(cfg, 0, n, dev), centre)
        out = pts.cpu().numpy()
        del pts
        return out

    # a small calibration pass sizes the sample: K + W passes must end within a few minutes
    n_cal = min(n_total, 4_000_000)
    cal = cpu_reference_run(cfg, first_points(n_cal), bmin, bmax, spacing, 1, 0, "calibration")
    n_sample = reference_sample_size(cfg, n_total, steps + warm, cal["value"])
    xyz_np = first_points(n_sample)
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    note = ("the full %d-point cloud" % n_total) if n_sample == n_total else (
        "first %d points of the %d-point cloud (bounded so that %d passes fit the time budget and host memory)"
        % (n_sample, n_total, steps + warm))
    r = cpu_reference_run(cfg, xyz_np, bmin, bmax, spacing, steps, warm, note)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["seconds_per_pass"] * 1e3,
            "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "f64/u64", "data": "synthetic",
            "config": config_block(cfg_name, cfg, n_total),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "points", "nodes")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default=None, choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--points", type=int, default=None, help="total points (default: the config's size)")
    ap.add_argument("--sort-mode", type=int, default=-1,
                    help="swgpu_set_sort_mode: -1 automatic (default), 0 eight LSD passes, 2 / 3 explicit top-digit sort")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed oracle check of the result")
    ap.add_argument("--no-payload", action="store_true", help="skip the attribute / writer-payload leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from schwarzwald_b200 import workloads
    cfg_name = args.config or workloads.default_config(max(world, args.gpus))
    cfg = workloads.CONFIGS[cfg_name]
    n_total = int(args.points) if args.points else cfg["points"]

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference_arm(args, cfg_name, cfg, n_total)

    import numpy as np
    import torch

    import schwarzwald_b200 as sw

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the tiler kernels have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # strong scaling: rank r generates points [r * n_total / world, (r + 1) * n_total / world) of the SAME cloud
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    n_local = hi - lo

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic input, resident in HBM before the timed region ---------------------------------
    mn, mx, xyz = full_cloud_tight_bounds(cfg, n_total, dev, start=lo, count=n_local, keep=(lo, n_local))
    if world > 1:
        t = torch.tensor(np.concatenate([mn, -mx]), device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        t = t.cpu().numpy()
        mn, mx = t[:3], -t[3:]
    bmin, bmax, spacing, centre = workloads.finish_bounds(cfg, mn, mx)
    xyz = workloads.apply_pre_transform(cfg, xyz, centre)
    sampling, tiling, conc = cfg["sampling"], cfg["tiling"], cfg["concurrency"]
    maxpts = workloads.MAX_POINTS_PER_NODE

    stream = torch.cuda.current_stream()
    if world > 1:
        from schwarzwald_b200.distributed import ShardedTiler
        tiler = ShardedTiler(sampling, tiling, bmin, bmax, spacing, max_points_per_node=maxpts, concurrency=conc,
                             device=local_rank)
    else:
        tiler = sw.GpuTiler(sampling, tiling, bmin, bmax, spacing, max_points_per_node=maxpts, concurrency=conc,
                            device=local_rank)
    tiler.set_stream(stream.cuda_stream)
    tiler.enable_timing(True)
    (tiler.tiler if world > 1 else tiler).set_sort_mode(args.sort_mode)

    ids_dev = None
    nodes_host = None

    def step():
        """index + sort + sampling sweep + reconstruct + hand-off (node table -> host, original ids on device)."""
        nonlocal ids_dev, nodes_host
        tiler.build_execution_graph(xyz)
        tiler.finalize()
        nn, ni = tiler.result_size()
        if ids_dev is None or ids_dev.numel() < ni:
            ids_dev = torch.empty(int(ni * 1.05) + 1024, dtype=torch.int32, device=dev)
        nodes_host = tiler.result_device_ids(ids_dev.data_ptr())

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
    sort_ms, finish_ms, launches = 0.0, 0.0, 0
    stats = None
    barrier()
    t_begin = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        stats = tiler.stats()
        sort_ms += stats["ms_sort"]
        finish_ms += stats["ms_sort_finish"]
        launches += stats["kernel_launches"]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    n_shard = n_local
    phase_ms = None
    if world > 1:  # one extra, untimed step with per-phase events (index / exchange / tile)
        tiler.profile = True
        step()
        phase_ms = {k: round(v, 3) for k, v in tiler.last["phase_ms"].items()}
        phase_ms["bytes_sent_off_gpu"] = tiler.last["bytes_sent_off_gpu"]
        phase_ms["exchange"] = tiler.last.get("exchange")
        n_shard = tiler.last["n_shard"]
        tiler.profile = False
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- payload leg: the same step with the attribute permutation north_star names ----------------------------
    # A 16-byte attribute record per point (LAS point format 2: 14 B, PointBuffer.h:291-304) travels with the point
    # (N > 1: through the exchange, 44 instead of 28 B per point), and every GPU produces the node-major writer
    # payloads of its nodes: the attribute records (swgpu_gather_attribute_device) and the float32 positions of
    # PNTSWriter (swgpu_get_payload_pnts_device).  Reported next to `value`, not inside it: the reference arm's
    # timed region hands over point ids only, and `value` keeps the same work as that arm.
    payload = None
    if not args.no_payload:
        try:
            nn0, ni0 = tiler.result_size()
            attr = torch.empty((max(n_local, 1), 16), dtype=torch.uint8, device=dev)
            attr.view(torch.int32).random_(0, 2 ** 31 - 1)
            cap = int(ni0 * 1.05) + 1024
            out_attr = torch.empty((cap, 16), dtype=torch.uint8, device=dev)
            out_pnts = torch.empty((cap, 3), dtype=torch.float32, device=dev)

            def step_payload():
                if world > 1:
                    tiler.build_execution_graph(xyz, attributes=attr[:n_local])
                else:
                    tiler.build_execution_graph(xyz)
                tiler.finalize()
                tiler.result_device_ids(ids_dev.data_ptr())
                if world > 1:
                    tiler.gather_attributes(out_attr)
                    tiler.tiler.payload_pnts_device(out_pnts.data_ptr())
                else:
                    tiler.gather_attribute_device(attr.data_ptr(), 16, out_attr.data_ptr())
                    tiler.payload_pnts_device(out_pnts.data_ptr())

            step_payload()
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p_steps = max(1, min(args.steps, 3))
            p0.record(stream)
            for _ in range(p_steps):
                step_payload()
            p1.record(stream)
            barrier()
            p_ms = p0.elapsed_time(p1) / p_steps
            if world > 1:
                t = torch.tensor([p_ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                p_ms = float(t.item())
            payload = {"ms_per_step": p_ms, "value": n_total / (p_ms * 1e-3), "unit": UNIT, "steps": p_steps,
                       "extra_ms_over_value_step": p_ms - ms_per_step,
                       "what": "the step of `value` plus: a 16-byte attribute record per point " +
                               ("through the exchange (44 instead of 28 B per point) and " if world > 1 else "") +
                               "gathered into node-major order, and the node-major float32 position payload "
                               "(PNTSWriter) of every node"}
            del attr, out_attr, out_pnts
        except Exception as ex:  # the headline numbers above stand on their own
            import traceback
            traceback.print_exc()
            payload = {"error": repr(ex)}
        torch.cuda.empty_cache()
        step()  # the parity and e2e legs below look at a result without attributes
        barrier()

    # ---- parity (untimed): this rank's result against the CPU oracle ----------------------------------------
    parity = None
    cpu_line = None
    n_nodes_local, n_ids_local = tiler.result_size()
    if not args.no_parity:
        try:
            from oracle import parity as par
            from oracle import sworacle
            full = world == 1 and n_total <= 100_000_000 and not args.no_cpu_baseline
            if full:
                # the cpu_baseline pass tiles the WHOLE cloud with the reference's code: its result is the checker
                r = cpu_reference_run(cfg, xyz.cpu().numpy(), bmin, bmax, spacing, 1, 0,
                                      "the full %d-point cloud" % n_total, keep_result=True)
                want, orc = r.pop("_result"), r.pop("_oracle")
                if want.duplicate_keys:  # std::sort leaves ties unspecified: re-tile with the pinned tie order
                    orc.set_reference_sort(False)
                    params = sworacle.make_params(sampling, tiling, spacing, bmin, bmax, max_points_per_node=maxpts,
                                                  concurrency=conc)
                    want = orc.tile(params, xyz.cpu().numpy(), copy=False, want_keys=False)
                cpu_line = r
                parity = par.full_parity(want, nodes_host, ids_dev[:n_ids_local])
                parity["start_level"] = [int(tiler.start_level()), int(want.start_level)]
                parity["ok"] = bool(parity["ok"] and tiler.start_level() == want.start_level)
                del want
            else:
                orc = sworacle.Oracle("ref" if sworacle.have_ref() else "port")
                orc.set_threads(max(1, len(os.sched_getaffinity(0)) // max(1, world)))
                if world > 1:
                    sx, sids = tiler.shard_positions()
                    depth = max(3, int(tiler.last.get("shard_levels", 3)))
                else:
                    sx, sids, depth = xyz, None, 3
                if sampling.startswith("MIN_DISTANCE") and tiling == "ACCURATE":
                    if sids is not None:
                        order = torch.argsort(sids.to(torch.int64) & 0xFFFFFFFF)
                        sorted_ids = (sids.to(torch.int64) & 0xFFFFFFFF)[order]

                        def xyz_of(idv):
                            q = torch.from_numpy(idv.astype(np.int64)).to(dev)
                            pos = torch.searchsorted(sorted_ids, q).clamp_(max=sorted_ids.numel() - 1)
                            return sx[order[pos]].cpu().numpy()
                    else:
                        def xyz_of(idv):
                            return sx[torch.from_numpy(idv.astype(np.int64)).to(dev)].cpu().numpy()
                    # nodes wholly inside this shard (spanning nodes are checked by tests/test_gpu_sharded.py on
                    # the merged result)
                    sl = int(tiler.last.get("shard_levels", 0)) if world > 1 else 0
                    local = nodes_host[nodes_host["levels"] >= sl]
                    parity = par.min_spacing_check(xyz_of, local, ids_dev[:n_ids_local], spacing)
                    if world > 1:  # nodes above the shard depth: the invariant on the node merged over all ranks
                        mine = par.spanning_node_parts(xyz_of, nodes_host, ids_dev[:n_ids_local], min(sl, 3))
                        everyone = [None] * world if rank == 0 else None
                        dist.gather_object(mine, everyone, dst=0)
                        if rank == 0:
                            parity["spanning_nodes"] = par.merged_min_spacing_check(everyone, spacing)
                            parity["ok"] = bool(parity["ok"] and parity["spanning_nodes"]["ok"])
                else:
                    parity = par.subtree_parity(orc, sampling, tiling, spacing, bmin, bmax, conc, maxpts, sx, sids,
                                                nodes_host, ids_dev[:n_ids_local], int(tiler.start_level()), depth=depth)
        except Exception as ex:  # a broken checker must not hide the measurement; it is reported as unchecked
            import traceback
            traceback.print_exc()
            parity = {"checked": False, "ok": False, "error": repr(ex)}
        if world > 1:
            flags = torch.tensor([1 if parity.get("checked") else 0, 1 if parity.get("ok") else 0,
                                  int(parity.get("nodes", 0) if isinstance(parity.get("nodes"), int) else 0),
                                  int(parity.get("ids", 0))], dtype=torch.int64, device=dev)
            gathered = [torch.zeros_like(flags) for _ in range(world)]
            dist.all_gather(gathered, flags)
            g = torch.stack(gathered).cpu().numpy()
            parity = {"checked": bool(g[:, 0].any()), "ok": bool(g[:, 1].all()), "ranks_checked": int(g[:, 0].sum()),
                      "nodes": int(g[:, 2].sum()), "ids": int(g[:, 3].sum()), "rank0": parity}

    # ---- e2e: host buffers in, node table + point ids out, copies inside the timed region -----------
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n_local, 3), dtype=torch.float64, pin_memory=True)
        host.copy_(xyz)
        nn, ni = tiler.result_size()
        ids_host = torch.empty(int(ni * 1.05) + 1024, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
        nodes_buf = np.empty(int(nn * 1.05) + 1024, sw.tiler.NODE_DTYPE)
        e_steps = max(1, min(args.steps, 3))
        t_e2e, d2h = [], 0
        if world == 1:
            host_np = host.numpy()
            for it in range(1 + e_steps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tiler.build_execution_graph(host_np)
                tiler.finalize()
                res = tiler.result(ids_out=ids_host, nodes_out=nodes_buf)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                if it > 0:
                    t_e2e.append(dt)
                d2h = int(res.ids.nbytes + res.nodes.nbytes)
            api = "swgpu_index_batch(host xyz) + swgpu_finalize + swgpu_get_nodes(host)"
        else:
            for it in range(1 + e_steps):
                barrier()
                t0 = time.perf_counter()
                xyz.copy_(host, non_blocking=True)  # this rank's slice: pinned host -> device
                tiler.build_execution_graph(xyz)
                tiler.finalize()
                res = tiler.result(ids_out=ids_host, nodes_out=nodes_buf)
                barrier()
                dt = time.perf_counter() - t0
                if it > 0:
                    t_e2e.append(dt)
                d2h = int(res.ids.nbytes + res.nodes.nbytes)
            api = ("per rank: cudaMemcpyAsync(pinned host slice) + ShardedTiler.build_execution_graph + finalize + "
                   "swgpu_get_nodes(host); wall clock between barriers")
        e_dt = sum(t_e2e) / len(t_e2e)
        h2d = int(n_local * 24)
        if world > 1:
            t = torch.tensor([e_dt, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            e_dt, h2d, d2h = float(tm[0].item()), int(t[1].item()), int(t[2].item())
        e2e = {"value": n_total / e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e_steps, "api": api}
        del host
        if world == 1 and cfg_name == "c2":
            # ---- the same call sequence fed with LAS record coordinates: 12 B/pt cross
            # PCIe instead of 24, the reader's int -> double conversion runs fused with the indexing kernel
            try:
                from schwarzwald_b200 import tiler as swt
                chunk = 1 << 24
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
                    res = tiler.result(ids_out=ids_host, nodes_out=nodes_buf)
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

    # per-rank stats -> rank 0
    total_bytes = stats["bytes_index"] + stats["bytes_sort"] + stats["bytes_gather"] + stats["bytes_sample"]
    agg = None
    if world > 1:
        t = torch.tensor([float(total_bytes), float(stats["n_nodes"]), float(stats["n_output_ids"]), float(n_shard),
                          float(stats["bytes_traffic"])], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        agg = {"sum": t.cpu().numpy(), "max": tmax.cpu().numpy()}
        bytes_all = float(agg["sum"][0])
    else:
        bytes_all = float(total_bytes)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    # dominant kernel: one onesweep radix pass: 12 B read + 12 B written per point of THIS GPU's shard.  The pass
    # count comes back through the library's byte accounting (8 + passes * 24 B per point).
    n_sorted = max(1, int(stats["n_points"]))
    n_passes = max(1, int(stats["sort_passes"]))
    pass_ms = (sort_ms - finish_ms) / args.steps / n_passes  # the segment finish kernel is not a pass
    sort_info = {"onesweep_passes": n_passes, "first_bit": int(stats["sort_first_bit"]),
                 "fallback": int(stats["sort_fallback"]), "ms_sort": sort_ms / args.steps,
                 "ms_segment_finish": finish_ms / args.steps,
                 "finish_scan_steps_per_point": stats["sort_scan_steps"] / n_sorted,
                 "finish_moved_fraction": stats["sort_moved"] / n_sorted,
                 "note": "passes over key bits >= first_bit, then runs of equal top bits are ordered in place by "
                         "segment_finish_kernel (same order as eight LSD passes; swgpu_set_sort_mode)"}
    achieved = (24.0 * n_sorted) / (pass_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "onesweep_pass_kernel (%d launches per step)" % n_passes,
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": captured_traffic(n_sorted), "algorithmic_bytes_per_launch": 24 * n_sorted,
                "whole_step": {"algorithmic_bytes": int(bytes_all),
                               "model": "SURVEY 8(d): index 36 + sort 8 + 8 passes x 24 (the model's pass count, whatever ran: see 'sort') + gather 2 x 24 (strategies that "
                                        "read positions) + per level (12 + 24 pos) r + 12 (r - s) + 4 s, all GPUs",
                               "achieved": bytes_all / (ms_per_step * 1e-3) / 1e9,
                               "frac": bytes_all / (ms_per_step * 1e-3) / 1e9 / (peak * world),
                               "traffic_model_bytes": int(stats["bytes_traffic"]) if world == 1 else int(agg["sum"][4])}}
    if world > 1 and phase_ms:
        ex_ms = sum(v for k, v in phase_ms.items() if isinstance(v, float) and ("partition" in k or "all_to_all" in k))
        off = float(phase_ms["bytes_sent_off_gpu"])
        if ex_ms > 0:
            roofline["nvlink"] = {"bytes_off_gpu": int(off), "exchange_ms": round(ex_ms, 3),
                                  "achieved": off / (ex_ms * 1e-3) / 1e9, "peak": NVLINK_PEAK_GBS, "unit": "GB/s",
                                  "frac": off / (ex_ms * 1e-3) / 1e9 / NVLINK_PEAK_GBS,
                                  "note": "rank 0: bytes this GPU writes into its peers' receive buffers / duration "
                                          "of the partition+exchange kernel; peak = NVLink 5 per direction per GPU"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64/u64", "data": "synthetic",
        "config": config_block(cfg_name, cfg, n_total),
        "run": {
            "points_per_gpu": int(n_local), "start_level": tiler.start_level(),
            "nodes": int(stats["n_nodes"]) if world == 1 else int(agg["sum"][1]),
            "output_ids": int(stats["n_output_ids"]) if world == 1 else int(agg["sum"][2]),
            "max_shard_points": int(n_shard) if world == 1 else int(agg["max"][3]),
            "step": "index + sort + sampling sweep + FAST reconstruct + hand-off (node table to host, node-major "
                    "original ids composed on the device)"},
        "roofline": roofline,
        "sort": sort_info,
        "stage_ms": {k: stats[k] for k in ("ms_index", "ms_sort", "ms_gather", "ms_sample", "ms_total")},
        "gpu_launches": int(launches),
        "shuffle_phase_ms": phase_ms,
        "clocks": clocks,
        "parity": parity,
        "parity_checked": bool(parity and parity.get("checked") and parity.get("ok")),
        "strong_scaling_base": (None if world == 1 else
                                {"note": "N > 1 runs strong-scale BASELINE configs[2..4]; the N = 1 bench line is configs[1] "
                                         "(another cloud and strategy), so value(N) / value(1) is NOT a scaling efficiency",
                                 "same_config_n1": same_config_n1(cfg_name)}),
    }
    if payload is not None:
        line["payload"] = payload
    if e2e is not None:
        line["e2e"] = e2e
    if cpu_line is None and not args.no_cpu_baseline and world == 1:
        # bounded sample of the same cloud (the full-size pass is the reference arm's job for big configs)
        n_s = min(n_total, 8_000_000 if cfg["sampling"].startswith("MIN_DISTANCE") else 30_000_000)
        cpu_line = cpu_reference_run(cfg, xyz[:n_s].cpu().numpy(), bmin, bmax, spacing, 1, 0,
                                     "first %d points of the %d-point cloud" % (n_s, n_total))
    if cpu_line is not None:
        line["cpu_baseline"] = {k: cpu_line[k] for k in ("value", "unit", "cores", "kind", "sample", "points", "nodes")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
