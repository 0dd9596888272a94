#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/all_v10.log
COUNTR_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 | tee gpurun_out/bench_pdl0.log
COUNTR_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 | tee gpurun_out/bench_pdl1.log
