#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches2.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
tail -1 gpurun_out/ncu_bench2.log | cut -c1-200
wc -l gpurun_out/launches2.csv
