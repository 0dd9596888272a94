#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/bench_n2.log | cut -c1-1500
if ! grep -q '"metric"' gpurun_out/bench_n2.log; then
  echo "--- retry eager"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > gpurun_out/bench_n2_eager.log 2>&1
  echo "rc=$?"; tail -5 gpurun_out/bench_n2_eager.log | cut -c1-1500
fi
