#!/bin/bash
mkdir -p gpurun_out
for grp in upsample2x_bwd gn_relu_backward colsum attention_backward cross_attn_core_bwd inorm_relu_pool_bwd conv_weight_grads decoder_gradients; do
  echo "=== $grp"
  timeout 300 python -m pytest tests/test_backward_gpu.py -q -m gpu -s -k "$grp" --no-header -p no:cacheprovider 2>&1 | tail -30 | tee "gpurun_out/bt_$grp.log" | tail -14
done
