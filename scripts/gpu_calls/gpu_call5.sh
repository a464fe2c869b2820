#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/c5_pytest.log | tail -30
for m in resnet50 wrn vgg16; do
  timeout 400 python bench.py --model $m --steps 20 --warmup 5 --repeats 5 > gpurun_out/c5_bench_$m.json 2> gpurun_out/c5_bench_$m.err
done
cat gpurun_out/c5_bench_*.json | cut -c1-330
for f in gpurun_out/c5_bench_*.err; do echo "== $f"; tail -5 $f | cut -c1-300; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1400 --csv --log-file gpurun_out/c5_launches_resnet50.csv \
     python bench.py --model resnet50 --steps 2 --warmup 3 --repeats 1 --no-graph > gpurun_out/c5_ncu_resnet50.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/c5_launches_wrn.csv \
     python bench.py --model wrn --steps 2 --warmup 3 --repeats 1 --no-graph > gpurun_out/c5_ncu_wrn.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/c5_launches_vgg16.csv \
     python bench.py --model vgg16 --steps 2 --warmup 3 --repeats 1 --no-graph > gpurun_out/c5_ncu_vgg16.log 2>&1
