#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm or conv3x3" --no-header -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/kt_gemm_v2.log
timeout 300 python -m pytest tests/test_backward_gpu.py -q -m gpu -s -k decoder_gradients --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/bt_decoder_gradients.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_v1.log
