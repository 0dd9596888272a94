#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trainer_gpu.py -q -m gpu -s --no-header -p no:cacheprovider 2>&1 | grep -E "finetuner|passed|failed|Error|assert|   \(" | tail -24 | tee gpurun_out/trainer.log
