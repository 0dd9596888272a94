#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/all_v9.log
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_v8.log | cut -c1-330
