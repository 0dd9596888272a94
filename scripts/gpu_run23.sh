#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_backward_gpu.py -q -m gpu -k "attention_backward" --no-header -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/attn_bwd.log
timeout 400 python -m pytest tests -q -m gpu --durations=6 --no-header -p no:cacheprovider 2>&1 | tail -14 | tee gpurun_out/all_v8.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/bench_v7.log
