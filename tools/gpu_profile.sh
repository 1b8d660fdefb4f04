#!/bin/bash
# ncu evidence for the bench step (1 GPU): launch list + full captures of the dominant kernels.
# usage: tools/gpu_profile.sh <tag>     outputs gpurun_out/<tag>_*
TAG=${1:-r01}
mkdir -p gpurun_out
KREGEX='regex:morton|digit_base|onesweep|gather_|node_rle|level_|level5|select_argmin|md_|compose_ids|key_histogram|tile_rank0|root_node|start_nodes|parent_nodes|partition|prefix_hist'
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 600 --csv \
  --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 10 -c 2 \
  -f -o gpurun_out/${TAG}_onesweep $BENCH > gpurun_out/${TAG}_onesweep.log 2>&1
echo "onesweep rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:level_count|level_scatter|morton" -c 5 \
  -f -o gpurun_out/${TAG}_sweep $BENCH > gpurun_out/${TAG}_sweep.log 2>&1
echo "sweep rc=$?"
