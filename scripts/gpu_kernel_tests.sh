#!/bin/bash
# Runs every kernel-test group in its own process (a trapped kernel poisons the CUDA context),
# each under a timeout so a protocol bug cannot hang the GPU box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in gemm_kmajor gemm_bf16 gemm_mn_major gemm_split gemm_batched gemm_epilogues "conv3x3 and not dx" conv3x3_dx layernorm attention_fwd cross_attn casts gn_relu exemplar; do
  name=$(echo "$grp" | tr ' ' '_')
  echo "=== $grp"
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -x --no-header -p no:cacheprovider 2>&1 | tail -25 | tee "gpurun_out/kt_$name.log" | tail -8
done
