#!/bin/bash
# End-of-iteration GPU pass: full parity suite, smoke, bench, IO kernels, ncu evidence.  usage: tools/gpu_round.sh <tag>
TAG=${1:-r01_v4}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
timeout 300 python tools/bench_io.py > gpurun_out/io_kernels.log 2>&1; echo "io rc=$?"; tail -30 gpurun_out/io_kernels.log
bash tools/gpu_profile.sh $TAG
