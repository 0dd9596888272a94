#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm or conv3x3" --durations=5 --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/kt_pair.log
timeout 300 python scripts/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm_pair.log
