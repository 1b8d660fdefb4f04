#!/bin/bash
# compute-sanitizer memcheck over small parity cases that touch every kernel family
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_memcheck.log \
  python -m pytest -x -q -m gpu \
  "tests/test_gpu_batches.py::test_multi_batch_matches_reference_golden_and_port[3]" \
  "tests/test_gpu_batches.py::test_multi_batch_matches_reference_golden_and_port[7]" \
  "tests/test_gpu_batches.py::test_multi_batch_terminal_nodes[JITTERED]" \
  "tests/test_gpu_parity.py::test_tiny_inputs" \
  "tests/test_gpu_multi_process_one.py::test_one_process_min_distance_invariant" \
  "tests/test_gpu_multi_process_one.py::test_one_process_attributes_and_reuse" \
  > gpurun_out/r02_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r02_memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r02_memcheck.log | sort | uniq -c | head -10
