#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trainer_gpu.py -q -m gpu -s --no-header -p no:cacheprovider 2>&1 | grep -E "finetuner|passed|failed|Error|assert" | tail -12 | tee gpurun_out/trainer.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v9_tuner.log | cut -c1-900
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --script-loop 2>&1 | tail -1 | tee gpurun_out/bench_v9_script.log | cut -c1-300
