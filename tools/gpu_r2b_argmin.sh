#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batches.py tests/test_gpu_sharded.py -x -q > gpurun_out/r02e_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r02e_tests.log
for c in "c3 100000000" "c1 10000000"; do
  set -- $c
  python bench.py --config $1 --points $2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-payload 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"$1\", round(d[\"ms_per_step\"],3), d[\"stage_ms\"], d[\"parity_checked\"], d[\"parity\"].get(\"ok\"))"
done
