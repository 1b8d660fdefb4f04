#!/bin/bash
# Round 2: single-pass level compaction: parity tests, then bench C2 / C3-100M with both compaction variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sort_modes.py -x -q > gpurun_out/r02c_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02c_tests.log
tail -4 gpurun_out/r02c_tests.log
for v in 1pass 2pass; do
  for c in "c2 100000000" "c3 100000000"; do
    set -- $c
    SWGPU_COMPACT=$v python bench.py --config $1 --points $2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-payload \
      > gpurun_out/r02c_bench_$1_$v.json 2> gpurun_out/r02c_bench_$1_$v.err
    python - <<PY
import json
for l in open("gpurun_out/r02c_bench_$1_$v.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$1 $v", round(d["ms_per_step"],3), d["stage_ms"], d["sort"]["onesweep_passes"], d["gpu_launches"])
PY
  done
done
