#!/bin/bash
# several configs at N GPUs, optionally reduced sizes.  usage: tools/gpu_r2_multi_cfgs.sh N "c4:200000000 c5:400000000 c3" [steps]
N=${1:-2}; LIST=${2:-"c3"}; K=${3:-2}
mkdir -p gpurun_out
for item in $LIST; do
  CFG=${item%%:*}; PTS=""; [[ "$item" == *:* ]] && PTS="--points ${item##*:}"
  TAG=r02_bench_${CFG}_n${N}${PTS:+_reduced}
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --config $CFG $PTS --steps $K --warmup 3 > gpurun_out/$TAG.log 2> gpurun_out/$TAG.err; echo "$TAG rc=$?"
  grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/$TAG.err | tail -8; cat gpurun_out/$TAG.log | cut -c1-6000
done
