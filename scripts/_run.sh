#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_kernel_tests.sh 2>&1 | grep -E "===|passed|failed" | tee gpurun_out/suite.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v11.log | cut -c1-260
