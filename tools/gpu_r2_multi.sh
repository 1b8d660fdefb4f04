#!/bin/bash
# Round-2 multi-GPU pass: NCCL sharded parity, then bench (default config c3, strong-scaled).  usage: tools/gpu_r2_multi.sh N [config]
N=${1:-2}; CFG=${2:-c3}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/sharded_parity.py > gpurun_out/r02_sharded_parity_n$N.log 2>&1; echo "parity rc=$?"
grep -E "PARITY|Error|error" gpurun_out/r02_sharded_parity_n$N.log | head -20
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --config $CFG --steps 3 --warmup 3 > gpurun_out/r02_bench_${CFG}_n$N.log 2> gpurun_out/r02_bench_${CFG}_n$N.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench_${CFG}_n$N.err; cat gpurun_out/r02_bench_${CFG}_n$N.log
