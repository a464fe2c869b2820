#!/bin/bash
# GPU call 6 (1 GPU): full suite; resnet/wrn/alexnet benches; sanitizer racecheck + synccheck; convergence bf16 vs tf32; N=1 BSP curve
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/c6_pytest.log | tail -20
for m in alexnet resnet50 wrn googlenet; do
  timeout 400 python bench.py --model $m --steps 20 --warmup 5 --repeats 5 > gpurun_out/c6_bench_$m.json 2> gpurun_out/c6_bench_$m.err
done
cat gpurun_out/c6_bench_*.json | cut -c1-330
timeout 300 python scripts/convergence.py --steps 320 > gpurun_out/c6_convergence_dtype.log 2>&1; grep CONVERGENCE gpurun_out/c6_convergence_dtype.log | cut -c1-1200
timeout 300 python scripts/convergence.py --steps 320 --bsp > gpurun_out/c6_convergence_n1.log 2>&1; grep CONVERGENCE gpurun_out/c6_convergence_n1.log | cut -c1-900
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 300 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tf32.py -m gpu -q -x -p no:cacheprovider \
    -k "(test_gemm_majors and 200-136-328) or (test_gemm_tf32_majors and 200-136-324) or (test_conv_fwd_bwd and cfg1) or (test_batch_norm and True-True) or test_rnn_ops or test_pool or dropout or softmax or (test_conv_group2 and 2-32-64)" \
    > gpurun_out/c6_sanitize_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/c6_sanitize_$tool.log; tail -6 gpurun_out/c6_sanitize_$tool.log
done
