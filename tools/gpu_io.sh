#!/bin/bash
# f2/f3 parity on the GPU + the whole GPU suite.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_io_transforms.py -m gpu -x -q > gpurun_out/pytest_io.log 2>&1; echo "io rc=$?"
tail -15 gpurun_out/pytest_io.log
