#!/bin/bash
mkdir -p gpurun_out
for grp in conv_weight_grads decoder_gradients; do
  timeout 300 python -m pytest tests/test_backward_gpu.py -q -m gpu -s -k "$grp" --no-header -p no:cacheprovider 2>&1 | tail -12 | tee "gpurun_out/bt_$grp.log"
done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -5 | tee gpurun_out/bench_graph.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_eager.log
