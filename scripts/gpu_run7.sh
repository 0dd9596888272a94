#!/bin/bash
mkdir -p gpurun_out
for w in proj qkv attn; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attention_fwd" -s 4 -c 1 -f -o gpurun_out/prof_$w python scripts/prof_gemm.py $w > gpurun_out/ncu_$w.log 2>&1
  tail -2 gpurun_out/ncu_$w.log
done
timeout 300 python -m pytest tests/test_backward_gpu.py -q -m gpu -s -k "decoder_gradients or exemplar_cnn" --no-header -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/bt_decoder_gradients.log
ls -la gpurun_out/*.ncu-rep
