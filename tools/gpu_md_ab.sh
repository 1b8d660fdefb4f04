#!/bin/bash
# MIN_DISTANCE wave kernels A/B: thread-per-cell threshold x number of ready queues (c4 at 125 M points, one GPU)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batches.py tests/test_gpu_sharded.py -q -k "MIN_DISTANCE or min_distance" 2>&1 | tail -3
for CFG in "0 1" "0 64" "6 1" "6 64" "6 16"; do
  set -- $CFG
  echo "== SWGPU_MD_THREAD_BELOW=$1 SWGPU_MD_QUEUES=$2"
  SWGPU_MD_THREAD_BELOW=$1 SWGPU_MD_QUEUES=$2 python bench.py --config c4 --points 125000000 --steps 2 --warmup 1 --no-e2e --no-parity --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']['ms_sample'])"
done
