#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -s --no-header -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/parity_fwd.log
timeout 600 python scripts/time_forward.py 2>&1 | tee gpurun_out/time_forward.log
