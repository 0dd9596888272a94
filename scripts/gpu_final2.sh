#!/bin/bash
# Last evidence pass of round 2 (one GPU): whole GPU suite, the three bench workloads, micro-benchmarks, launch list, stage times.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_tests_final.log; cat gpurun_out/r2_tests_final.log
timeout 100 python scripts/bench_gn_bwd.py > gpurun_out/r2_bench_gn_bwd.log 2>&1; cat gpurun_out/r2_bench_gn_bwd.log
COUNTR_GN_DOT_STAGED=0 timeout 100 python scripts/bench_gn_bwd.py 2>&1 | grep conv1x1
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; cut -c1-200 gpurun_out/r2_bench_final.json
timeout 300 python bench.py --workload infer0 --no-cpu-baseline > gpurun_out/r2_bench_infer0.json 2> gpurun_out/r2_bench_infer0.err; cut -c1-200 gpurun_out/r2_bench_infer0.json
timeout 300 python bench.py --workload pretrain --no-cpu-baseline > gpurun_out/r2_bench_pretrain.json 2> gpurun_out/r2_bench_pretrain.err; cut -c1-200 gpurun_out/r2_bench_pretrain.json
bash scripts/gpu_launch_list.sh r2_launches_final > /dev/null 2>&1
python scripts/time_stages.py 2>&1 | tail -2 > gpurun_out/r2_time_stages.log; cat gpurun_out/r2_time_stages.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_backward_gpu.py tests/test_kernels_gpu.py tests/test_augment_gpu.py -x -q -k "gn_relu or weight_refresh or attention_fwd or mosaic or affine" 2>&1 | tail -6 > gpurun_out/r2_sanitizer_memcheck_late.log; cat gpurun_out/r2_sanitizer_memcheck_late.log
STEP="python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline --no-extras"
cap() {   # name, kernel regex, launch skip
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count 1 -f -o gpurun_out/$1 $STEP > /dev/null 2>&1
  python scripts/ncu_summary.py gpurun_out/$1.ncu-rep gpurun_out/$1.json > /dev/null 2>&1 && echo "captured $1"
}
cap r2_ncu_gn_head_reduce gn_head_reduce_kernel 2
cap r2_ncu_gn_up2_staged gn_relu_up2_staged_kernel 4
rm -f gpurun_out/*.ncu-rep
