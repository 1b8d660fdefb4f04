#!/bin/bash
# Round 2 (second half): launch lists with the top-digit sort, ncu --set full of the sort kernels
mkdir -p gpurun_out
bash tools/gpu_launchlist_cfg.sh c2 100000000 r02b_c2 "--sort-mode 3" > gpurun_out/r02b_c2_summary.txt 2>&1
tail -30 gpurun_out/r02b_c2_summary.txt
SWGPU_SORT_FIRST_PASS=3 timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name "regex:^(onesweep_pass_kernel|segment_finish_kernel)" --launch-skip 3 -c 3 -o gpurun_out/r02b_sort -f \
  python bench.py --config c2 --steps 1 --warmup 0 --no-e2e --no-parity --no-cpu-baseline --no-payload > gpurun_out/r02b_sort.log 2>&1
echo "ncu rc=$?"
