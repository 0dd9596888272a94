#!/bin/bash
# launch list of one eager fine-tune step (cold-cache, serialised: compare shares, not absolutes)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_backward_gpu.py -q -m gpu -s -k decoder_gradients --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/bt_decoder_gradients.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
wc -l gpurun_out/launches.csv
