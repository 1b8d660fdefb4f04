#!/bin/bash
# compute-sanitizer over the kernels of the second half of round 2: racecheck (shared memory) and memcheck on the sort-mode
# tests, the single-pass compaction variant and a tiling case per sort mode
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/r02b_racecheck.log \
  python -m pytest -x -q -m gpu tests/test_gpu_sort_modes.py -k "one_by_one or limit or ties or ragged" \
  "tests/test_gpu_parity.py::test_single_pass_compaction_variant[RANDOM_GRID-FAST]" \
  > gpurun_out/r02b_racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r02b_racecheck_pytest.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/r02b_racecheck.log | sort | uniq -c | sort -rn | head -10
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02b_memcheck.log \
  python -m pytest -x -q -m gpu tests/test_gpu_sort_modes.py -k "one_by_one or long or every_sort_mode" \
  "tests/test_gpu_parity.py::test_single_pass_compaction_variant[JITTERED-ACCURATE]" \
  > gpurun_out/r02b_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r02b_memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r02b_memcheck.log | sort | uniq -c | head -10
