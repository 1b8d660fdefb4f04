#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over small cases of the kernels rewritten in round 2
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/r02_racecheck.log \
  python -m pytest -x -q -m gpu \
  "tests/test_gpu_batches.py::test_multi_batch_matches_reference_golden_and_port[1]" \
  "tests/test_gpu_batches.py::test_multi_batch_matches_reference_golden_and_port[8]" \
  "tests/test_gpu_parity.py::test_tiny_inputs" \
  > gpurun_out/r02_racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r02_racecheck_pytest.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/r02_racecheck.log | sort | uniq -c | sort -rn | head -10
