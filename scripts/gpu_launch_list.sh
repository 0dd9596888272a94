#!/bin/bash
# ncu launch list of one eager fine-tune step (same command as the bench): gpurun_out/$1.csv
name=${1:-r2_launches}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/$name.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/$name.log 2>&1
tail -1 gpurun_out/$name.log | cut -c1-200
wc -l gpurun_out/$name.csv
