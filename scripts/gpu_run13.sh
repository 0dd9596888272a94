#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_noct_gpu.py -q -m gpu -s --no-header -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/noct.log
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -3
