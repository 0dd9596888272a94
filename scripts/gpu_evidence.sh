#!/bin/bash
# Round-2 evidence run (one GPU): compute-sanitizer logs and `ncu --set full` captures of the kernels VERDICT r1 asked for.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_evidence.sh'   ->  gpurun_out/r2_*  (summaries are copied into profiles/ by hand)
mkdir -p gpurun_out
SEL='gemm_linear or gemm_epilogues or gemm_aux or grouped or attention_fwd or conv3x3 or layernorm'
echo "=== compute-sanitizer memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 --log-file gpurun_out/r2_sanitizer_memcheck.log \
   python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" -p no:cacheprovider 2>&1 | tail -3
tail -4 gpurun_out/r2_sanitizer_memcheck.log
echo "=== compute-sanitizer racecheck"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 --log-file gpurun_out/r2_sanitizer_racecheck.log \
   python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm_epilogues or grouped or attention_fwd_growing or layernorm" -p no:cacheprovider 2>&1 | tail -3
tail -4 gpurun_out/r2_sanitizer_racecheck.log
STEP="python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline --no-extras"
cap() {   # name, kernel regex, launch skip
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count 1 -f -o gpurun_out/$1 $STEP > /dev/null 2>&1
  python scripts/ncu_summary.py gpurun_out/$1.ncu-rep gpurun_out/$1.json > /dev/null 2>&1 && echo "captured $1"
}
# skip counts: 3 steps run (2 warm-up + 1), take launches of the LAST step
cap r2_ncu_attn_bwd32 attention_bwd_kernel 4
cap r2_ncu_gn_bwd_reduce0 "gn_relu_bwd_reduce_kernel<0>" 6
cap r2_ncu_gn_bwd_reduce1 "gn_relu_bwd_reduce_kernel<1>" 2
cap r2_ncu_gn_bwd_apply gn_bwd_apply_kernel 8
cap r2_ncu_layernorm_bwd layernorm_bwd_kernel 14
cap r2_ncu_weight_refresh weight_refresh_kernel 2
cap r2_ncu_grouped_dw grouped_dw_kernel 2
cap r2_ncu_attn_fwd64 "attention_fwd_kernel<64" 24
cap r2_ncu_gn_relu_up2 gn_relu_up2_rows_kernel 8
cap r2_ncu_conv_h3 "gemm_kernel<\(bool\)1>" 22
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2_ncu_lin_qkv python scripts/prof_linear.py qkv > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r2_ncu_lin_qkv.ncu-rep gpurun_out/r2_ncu_lin_qkv.json > /dev/null 2>&1 && echo "captured qkv"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2_ncu_lin_fc2 python scripts/prof_linear.py fc2 > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r2_ncu_lin_fc2.ncu-rep gpurun_out/r2_ncu_lin_fc2.json > /dev/null 2>&1 && echo "captured fc2"
ls -la gpurun_out/r2_ncu_*.json | wc -l
