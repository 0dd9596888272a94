#!/bin/bash
# Round-2 final evidence (one GPU): whole GPU suite, the three bench workloads, launch lists, the ncu captures the first pass missed.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_tests_final.log; cat gpurun_out/r2_tests_final.log
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; cut -c1-200 gpurun_out/r2_bench_final.json
timeout 300 python bench.py --workload infer0 --no-cpu-baseline > gpurun_out/r2_bench_infer0.json 2> gpurun_out/r2_bench_infer0.err; cut -c1-200 gpurun_out/r2_bench_infer0.json
timeout 300 python bench.py --workload pretrain --no-cpu-baseline > gpurun_out/r2_bench_pretrain.json 2> gpurun_out/r2_bench_pretrain.err; cut -c1-200 gpurun_out/r2_bench_pretrain.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_reference_arm.json
bash scripts/gpu_launch_list.sh r2_launches_final > /dev/null 2>&1
if [ -n "$FULL" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2_launches_pretrain_final.csv python bench.py --workload pretrain --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > /dev/null 2>&1
python scripts/cublas_compare.py > gpurun_out/r2_cublas_final.log 2>&1; tail -10 gpurun_out/r2_cublas_final.log
fi
timeout 100 python scripts/bench_gn_bwd.py > gpurun_out/r2_bench_gn_bwd.log 2>&1; cat gpurun_out/r2_bench_gn_bwd.log
timeout 100 python scripts/bench_attn.py > gpurun_out/r2_bench_attn.log 2>&1; cat gpurun_out/r2_bench_attn.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_backward_gpu.py tests/test_kernels_gpu.py tests/test_augment_gpu.py -x -q -k "gn_relu or weight_refresh or attention_fwd or mosaic or affine" 2>&1 | tail -6 > gpurun_out/r2_sanitizer_memcheck_late.log; cat gpurun_out/r2_sanitizer_memcheck_late.log
python scripts/time_stages.py 2>&1 | tail -2 > gpurun_out/r2_time_stages.log; cat gpurun_out/r2_time_stages.log
STEP="python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline --no-extras"
cap() {   # name, kernel regex, launch skip
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count 1 -f -o gpurun_out/$1 $STEP > /dev/null 2>&1
  python scripts/ncu_summary.py gpurun_out/$1.ncu-rep gpurun_out/$1.json > /dev/null 2>&1 && echo "captured $1"
}
cap r2_ncu_gn_bwd_reduce1 gn_relu_bwd_reduce_kernel 2
cap r2_ncu_gn_bwd_gather gn_relu_bwd_gather 6
cap r2_ncu_weight_refresh weight_refresh_kernel 2
if [ -n "$FULL" ]; then
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_conv_h3 python scripts/prof_gemm.py conv > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r2_conv_h3.ncu-rep gpurun_out/r2_conv_h3_ncu_summary.json > /dev/null 2>&1 && echo "captured conv h3"
fi
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attention_fwd4 --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_ncu_attn_fwd4 python scripts/prof_gemm.py attn > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r2_ncu_attn_fwd4.ncu-rep gpurun_out/r2_ncu_attn_fwd4.json > /dev/null 2>&1 && echo "captured attn fwd4"
rm -f gpurun_out/*.ncu-rep
