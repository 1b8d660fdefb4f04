#!/bin/bash
# Round-2 first GPU pass: parity suite (incl. full-size), bench c2 / c3 on one GPU, reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_nvidia_smi.txt 2>&1
free -g > gpurun_out/r02_host.txt; nproc >> gpurun_out/r02_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c2.log 2> gpurun_out/r02_bench_c2.err; echo "bench c2 rc=$?"
tail -5 gpurun_out/r02_bench_c2.err; cat gpurun_out/r02_bench_c2.log
timeout 900 python bench.py --config c3 --steps 3 --warmup 3 > gpurun_out/r02_bench_c3_n1.log 2> gpurun_out/r02_bench_c3_n1.err; echo "bench c3 rc=$?"
tail -5 gpurun_out/r02_bench_c3_n1.err; cat gpurun_out/r02_bench_c3_n1.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r02_ref_c2.log 2> gpurun_out/r02_ref_c2.err; echo "ref c2 rc=$?"
tail -5 gpurun_out/r02_ref_c2.err; cat gpurun_out/r02_ref_c2.log
