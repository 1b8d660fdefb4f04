#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02d_tests.log 2>&1
echo "tests exit $?"; tail -2 gpurun_out/r02d_tests.log
for c in "c2 100000000" "c3 100000000"; do
  set -- $c
  python bench.py --config $1 --points $2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-payload \
      > gpurun_out/r02d_bench_$1.json 2> gpurun_out/r02d_bench_$1.err
  python - <<PY
import json
for l in open("gpurun_out/r02d_bench_$1.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$1", round(d["ms_per_step"],3), d["stage_ms"], d["sort"]["onesweep_passes"], d["gpu_launches"], round(d["roofline"]["whole_step"]["frac"],3))
PY
done
SWGPU_SORT_FIRST_PASS=3 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name "regex:^(level_count_kernel|level_scatter_kernel|level_scan_kernel)" -c 3 \
  -o gpurun_out/r02d_sweep -f python bench.py --config c2 --steps 1 --warmup 0 --no-e2e --no-parity --no-cpu-baseline --no-payload > gpurun_out/r02d_sweep.log 2>&1
echo "ncu rc=$?"
