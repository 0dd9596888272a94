#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/all_v7.log
timeout 300 python scripts/bench_gemm.py 2>&1 | tee gpurun_out/bench_elem.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/bench_v6.log
