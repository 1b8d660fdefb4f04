#!/bin/bash
# MIN_DISTANCE: parity subset, then timing probes with hard timeouts.  usage: tools/gpu_md.sh [sizes...]
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "MIN_DISTANCE or min_distance" > gpurun_out/pytest_md.log 2>&1; echo "md parity rc=$?"
tail -4 gpurun_out/pytest_md.log
for n in "$@"; do
  timeout 120 python tools/bench_matrix.py --points $n --steps 1 \
    --cases c3_terrain_min_distance_accurate,x_urban_min_distance_fast,x_terrain_min_distance_fast_fast \
    --out gpurun_out/md_$n.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d.get(k) for k in ('case','points','wall_ms','ms_sample','n_levels','min_distance_rounds','points_per_s','error')})
    else: print(l, end='')
"
  echo "n=$n rc=$?"
done
