#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trainer_gpu.py -q -m gpu -s --no-header -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/trainer.log
