#!/bin/bash
# ncu --set full of selected kernels of one bench step.  usage: tools/gpu_ncu_full_cfg.sh CFG POINTS TAG "k1|k2|..." [launch-skip] [count]
CFG=${1:-c3}; PTS=${2:-100000000}; TAG=${3:-r02_full}; KERN=${4:-select_argmin_kernel}; SKIP=${5:-0}; CNT=${6:-4}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name "regex:^($KERN)" --launch-skip $SKIP -c $CNT \
  -o gpurun_out/${TAG} -f python bench.py --config $CFG --points $PTS --steps 1 --warmup 0 --no-e2e --no-parity --no-cpu-baseline --no-payload \
  > gpurun_out/${TAG}.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}.log | cut -c1-300
