#!/usr/bin/env python
"""Device-side timing of the f2/f3 kernels: the fused LAS-record indexing kernel and the two
writer-payload kernels, on the bench workload (100 M terrain points), against their algorithmic bytes.

    python tools/bench_io.py [--points 100000000] [--out gpurun_out/io_kernels.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=100_000_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "io_kernels.json"))
    args = ap.parse_args()
    import numpy as np
    import torch

    import schwarzwald_b200 as sw
    from schwarzwald_b200 import synth
    from schwarzwald_b200 import tiler as swt

    dev = torch.device("cuda", 0)
    n = args.points
    las = torch.empty((n, 3), dtype=torch.int32, device=dev)
    chunk = 1 << 24
    mn = mx = None
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        x = synth.terrain(m, seed=2, device=dev, start=s)
        lo, hi = synth.tight_bounds(x)
        mn = lo if mn is None else np.minimum(mn, lo)
        mx = hi if mx is None else np.maximum(mx, hi)
        las[s:s + m] = torch.round((x - torch.tensor(np.floor(lo), device=dev)) / 0.001).to(torch.int32)
    offset = np.floor(mn)
    # every chunk used its own floor(min) above only if the chunks differ; recompute with the global offset
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        x = synth.terrain(m, seed=2, device=dev, start=s)
        las[s:s + m] = torch.round((x - torch.tensor(offset, device=dev)) / 0.001).to(torch.int32)
    del x
    bmin, bmax = sw.cubic_bounds(mn, mx)
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    tr = swt.las_transform([0.001] * 3, offset, mn - 1.0, mx + 1.0)
    peak = 6549.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    stream = torch.cuda.current_stream()
    rows = {}
    with sw.GpuTiler("RANDOM_GRID", "FAST", bmin, bmax, spacing, concurrency=32) as t:
        t.set_stream(stream.cuda_stream)
        t.enable_timing(True)
        ms = []
        for it in range(4):
            t.build_execution_graph_las(las, tr)
            t.finalize()
            torch.cuda.synchronize()
            if it:
                ms.append(t.stats()["ms_index"])
        ms_index = sum(ms) / len(ms)
        rows["las_encode_kernel"] = {"ms": ms_index, "algorithmic_bytes": 44 * n,
                                     "GBps": 44 * n / ms_index / 1e6, "frac_of_measured_peak": 44 * n / ms_index / 1e6 / peak}
        _, ni = t.result_size()
        out_f = torch.empty((ni, 3), dtype=torch.float32, device=dev)
        out_i = torch.empty((ni, 3), dtype=torch.int32, device=dev)
        for name, call, buf in (("payload_pnts_kernel", t.payload_pnts_device, out_f),
                                ("payload_las_kernel", t.payload_las_device, out_i)):
            times = []
            for it in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                call(buf.data_ptr())
                e1.record(stream)
                torch.cuda.synchronize()
                if it:
                    times.append(e0.elapsed_time(e1))
            msk = sum(times) / len(times)
            alg = (4 + 4 + 24 + 12) * ni  # node-major index, permutation entry, position gather, payload record
            rows[name] = {"ms": msk, "records": ni, "algorithmic_bytes": alg, "GBps": alg / msk / 1e6,
                          "frac_of_measured_peak": alg / msk / 1e6 / peak,
                          "note": "includes the compose of nothing else; the LAS variant also uploads the node headers "
                                  "and synchronises (host-side header table)" if "las" in name else ""}
    print(json.dumps(rows, indent=1))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
