#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/time_torch_eager.py 2>&1 | tail -3 | tee gpurun_out/torch_eager.log
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_backward_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v5.log
