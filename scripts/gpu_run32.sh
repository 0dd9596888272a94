#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_infer_gpu.py tests/test_parity_gpu.py -q -m gpu -s --no-header -p no:cacheprovider 2>&1 | grep -E "parity large|passed|failed|Error|assert" | tail -12 | tee gpurun_out/infer.log
timeout 300 python scripts/time_inference.py 2>&1 | tail -4 | tee gpurun_out/time_inference.log
