#!/bin/bash
# A/B of library builds: quick parity subset + bench stage times for each SWGPU_LIB given.
mkdir -p gpurun_out
for lib in "$@"; do
  export SWGPU_LIB=$PWD/schwarzwald_b200/$lib
  echo "=== $lib"
  timeout 600 python -m pytest tests -m gpu -x -q -k "primitives or terrain_medium or tiny or outliers" 2>&1 | tail -2
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(json.dumps({'ms': d['ms_per_step'], 'stage': d['stage_ms'], 'roof': d['roofline']['frac']}))
    else: print(l, end='')
"
done
