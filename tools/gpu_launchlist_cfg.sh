#!/bin/bash
# ncu launch list (gpu__time_duration only) of one bench step of a config, library kernels only.
# usage: tools/gpu_launchlist_cfg.sh CFG POINTS TAG [extra bench.py arguments, e.g. "--sort-mode 3"]
CFG=${1:-c3}; PTS=${2:-100000000}; TAG=${3:-r02_${CFG}}; EXTRA=${4:-}
mkdir -p gpurun_out
REGEX='regex:^(morton|onesweep|gather|level|select|argmin|node_|tile_rank|start_nodes|parent|md_|compose|las_|payload|partition|prefix|store_|scan_|merge_|concat_|face_|bin_|root_node|make_gids|key_hist|sort_|segment_finish|long_run|run_stats|digit_base)'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$REGEX" -c 3000 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --config $CFG --points $PTS --steps 1 --warmup 1 \
  --no-e2e --no-parity --no-cpu-baseline --no-payload $EXTRA > gpurun_out/${TAG}_launches.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_launches.log | cut -c1-600
python profiles/summarize_launches.py gpurun_out/${TAG}_launches.csv
