#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/prof_conv_h3 python scripts/prof_gemm.py conv > gpurun_out/ncu_conv.log 2>&1
tail -1 gpurun_out/ncu_conv.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_bwd -s 1 -c 1 -f -o gpurun_out/prof_attn_bwd python -m pytest tests/test_backward_gpu.py -q -m gpu -k "attention_backward and 576-16-32-True" -p no:cacheprovider > gpurun_out/ncu_attn_bwd.log 2>&1
tail -1 gpurun_out/ncu_attn_bwd.log
bash scripts/gpu_profile2.sh
