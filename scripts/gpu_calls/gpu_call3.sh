#!/bin/bash
# GPU call 3 (1 GPU): full GPU suite (tf32 MN-major fix, BN kernels, native ResNet / WRN) + benches of the native residual nets
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
for m in resnet50 wrn; do
  timeout 400 python bench.py --model $m --steps 20 --warmup 5 --repeats 5 > gpurun_out/c3_bench_$m.json 2> gpurun_out/c3_bench_$m.err
done
timeout 300 python bench.py --model alexnet --dtype tf32 --steps 20 --warmup 5 > gpurun_out/c3_bench_alexnet_tf32.json 2> gpurun_out/c3_bench_alexnet_tf32.err
grep -E "passed|failed|FAILED|rc=" gpurun_out/c3_pytest.log | tail -40
cat gpurun_out/c3_bench_*.json | cut -c1-330
for f in gpurun_out/c3_*.err; do echo "== $f"; tail -6 $f | cut -c1-300; done
