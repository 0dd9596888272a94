#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/all_v5.log
timeout 300 python scripts/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm_v5.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v4.log
