#!/bin/bash
# ncu launch lists (gpu__time_duration) of one tiling step per strategy.  usage: tools/gpu_matrix_launches.sh <tag> case...
TAG=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${TAG}_${c}_launches.csv python tools/bench_matrix.py --points 100000000 --steps 1 --cases $c \
    --out gpurun_out/${TAG}_${c}.json > gpurun_out/${TAG}_${c}.log 2>&1
  echo "$c rc=$?"
done
