#!/bin/bash
# A/B: L2 fetch granularity for the random gathers (c3 at 100 M points, stage times of the bench line)
for G in 0 32 64; do
  echo "== SWGPU_L2_FETCH=$G"
  SWGPU_L2_FETCH=$G python bench.py --config c3 --points 100000000 --steps 3 --warmup 2 --no-e2e --no-parity --no-cpu-baseline --no-payload 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
  SWGPU_L2_FETCH=$G python bench.py --config c2 --steps 3 --warmup 2 --no-e2e --no-parity --no-cpu-baseline --no-payload 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
done
