#!/bin/bash
# N-GPU run: multi-rank parity over NCCL, then the bench at N GPUs.  usage: tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/sharded_parity.py > gpurun_out/sharded_parity_n$N.log 2>&1; echo "parity rc=$?"
grep -E "PARITY|Error|error" gpurun_out/sharded_parity_n$N.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.log | cut -c1-1800
