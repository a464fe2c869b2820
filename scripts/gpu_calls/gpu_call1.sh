#!/bin/bash
# GPU call 1 (1 GPU): regression tests + every model through the new bench + the library yardstick
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
for m in alexnet googlenet vgg16 resnet50 wrn; do
  timeout 300 python bench.py --model $m --steps 20 --warmup 5 --repeats 5 > gpurun_out/c1_bench_$m.json 2> gpurun_out/c1_bench_$m.err
done
for m in alexnet googlenet vgg16 resnet50 wrn; do
  timeout 300 python bench.py --impl torch_best --model $m --steps 20 --warmup 5 --repeats 5 > gpurun_out/c1_torch_$m.json 2> gpurun_out/c1_torch_$m.err
done
timeout 300 python bench.py --impl torch_best --dtype tf32 --model alexnet --steps 20 --warmup 5 > gpurun_out/c1_torch_alexnet_tf32.json 2> gpurun_out/c1_torch_alexnet_tf32.err
tail -3 gpurun_out/c1_pytest.log
cat gpurun_out/c1_bench_*.json gpurun_out/c1_torch_*.json | cut -c1-400
tail -5 gpurun_out/c1_*.err
