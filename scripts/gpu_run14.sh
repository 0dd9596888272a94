#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_n2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/prof_conv_h3 python scripts/prof_gemm.py conv > gpurun_out/ncu_conv.log 2>&1
tail -1 gpurun_out/ncu_conv.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.log
