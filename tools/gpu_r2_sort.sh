#!/bin/bash
# Round 2, sort: the sort-mode tests, a memcheck pass over a few of them, bench C2 in every sort mode.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sort_modes.py -q > gpurun_out/r02s_sort_tests.log 2>&1
echo "sort tests exit $?" >> gpurun_out/r02s_sort_tests.log
tail -5 gpurun_out/r02s_sort_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sort_modes.py -x -q \
  -k "limit or one_inversion or ragged or one_by_one" > gpurun_out/r02s_sort_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02s_sort_memcheck.log
tail -4 gpurun_out/r02s_sort_memcheck.log
for m in -1 3; do
  python bench.py --config c2 --steps 5 --warmup 3 --sort-mode $m --no-cpu-baseline --no-e2e --no-parity --no-payload \
    > gpurun_out/r02s_bench_c2_mode$m.json 2> gpurun_out/r02s_bench_c2_mode$m.err
  python - <<PY
import json
for l in open("gpurun_out/r02s_bench_c2_mode$m.json"):
    if l.startswith("{"):
        d = json.loads(l); print("mode $m", d["ms_per_step"], d["sort"], d["roofline"]["frac"], d["stage_ms"])
PY
done
python bench.py --config c3 --points 100000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-payload \
    > gpurun_out/r02s_bench_c3_100m.json 2> gpurun_out/r02s_bench_c3_100m.err
tail -c 1500 gpurun_out/r02s_bench_c3_100m.json
