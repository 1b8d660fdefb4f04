#!/bin/bash
# ncu --set full of segment_finish_kernel on C2 for first_pass = 3 and 2
mkdir -p gpurun_out
for m in 3 2; do
SWGPU_SORT_FIRST_PASS=$m timeout 900 ncu --set full --clock-control none --import-source on --kernel-name "regex:^(segment_finish_kernel)" -c 1 \
  -o gpurun_out/r02s_finish_m$m -f python bench.py --config c2 --steps 1 --warmup 0 --no-e2e --no-parity --no-cpu-baseline --no-payload \
  > gpurun_out/r02s_finish_m$m.log 2>&1
echo "ncu rc=$?"
done
