#!/bin/bash
# launch list of one bench step.  usage: tools/gpu_launchlist.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
KREGEX='regex:morton|digit_base|onesweep|gather_|node_rle|level_|level5|select_argmin|md_|compose_ids|key_histogram|tile_rank0|root_node|partition|prefix_hist'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 600 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
