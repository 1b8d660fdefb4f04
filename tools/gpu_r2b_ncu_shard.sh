#!/bin/bash
# ncu --set full of the partition kernels (one GPU, local destinations) and the payload kernels
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --kernel-name "regex:^(partition_scatter_kernel|partition_count_kernel)" --launch-skip 2 -c 2 \
  -o gpurun_out/r02b_partition -f python tools/bench_shard_kernels.py 50000000 > gpurun_out/r02b_partition.log 2>&1
echo "ncu partition rc=$?"
