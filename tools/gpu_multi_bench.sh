#!/bin/bash
# bench.py at N GPUs (torchrun).  usage: tools/gpu_multi_bench.sh N [steps] [warmup]
N=${1:-2}; K=${2:-3}; W=${3:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps $K --warmup $W > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.log | cut -c1-2500
