#!/bin/bash
# MIN_DISTANCE: coarser cells on sparse levels (c4 at 125 M points, one GPU) + parity
for CFG in "0 6" "1 6" "1 12" "2 12" "1 24" "2 24"; do
  set -- $CFG
  echo "== SWGPU_MD_COARSEN=$1 SWGPU_MD_COARSEN_BELOW=$2"
  SWGPU_MD_COARSEN=$1 SWGPU_MD_COARSEN_BELOW=$2 python bench.py --config c4 --points 125000000 --steps 2 --warmup 1 --no-e2e --no-parity --no-cpu-baseline --no-payload 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']['ms_sample'])"
done
SWGPU_MD_COARSEN=1 SWGPU_MD_COARSEN_BELOW=12 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batches.py tests/test_gpu_sharded.py -q -k "MIN_DISTANCE or min_distance" 2>&1 | tail -3
