#!/bin/bash
# Whole GPU suite, one pytest process per file (a trapped kernel poisons the CUDA context of its process only),
# each under a timeout so a protocol bug cannot hang the GPU box.  Usage: gpurun -- 'bash scripts/gpu_kernel_tests.sh'
mkdir -p gpurun_out
for f in tests/test_kernels_gpu.py tests/test_backward_gpu.py tests/test_parity_gpu.py tests/test_noct_gpu.py tests/test_trainer_gpu.py tests/test_infer_gpu.py; do
  echo "=== $f"
  timeout 300 python -m pytest "$f" -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -6 | tee "gpurun_out/$(basename "$f" .py).log" | tail -3
done
