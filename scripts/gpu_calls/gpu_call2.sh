#!/bin/bash
# GPU call 2 (1 GPU): full GPU test-suite incl. the tf32 mode; AlexNet / VGG16 / GoogLeNet in tf32
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
for m in alexnet vgg16 googlenet; do
  timeout 300 python bench.py --model $m --dtype tf32 --steps 20 --warmup 5 --repeats 5 > gpurun_out/c2_bench_${m}_tf32.json 2> gpurun_out/c2_bench_${m}_tf32.err
done
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/c2_bench_alexnet_bf16.json 2> gpurun_out/c2_bench_alexnet_bf16.err
tail -15 gpurun_out/c2_pytest.log
cat gpurun_out/c2_bench_*.json | cut -c1-300
for f in gpurun_out/c2_*.err; do echo "== $f"; tail -5 $f; done
