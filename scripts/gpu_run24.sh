#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dw_cta_pair or cta_pair" --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/kt_pair_dw.log
timeout 300 python scripts/bench_gemm.py 2>&1 | tee gpurun_out/bench_dw_pair.log
