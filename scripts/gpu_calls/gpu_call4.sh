#!/bin/bash
# GPU call 4 (1 GPU): full GPU suite; epilogue probe; googlenet / resnet50 benches; per-launch time lists; ncu captures
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/c4_pytest.log | tail -30
timeout 300 python scripts/epilogue_probe.py > gpurun_out/c4_epilogue_probe.txt 2>&1; cat gpurun_out/c4_epilogue_probe.txt
for m in googlenet resnet50 alexnet; do
  timeout 400 python bench.py --model $m --steps 20 --warmup 5 --repeats 5 > gpurun_out/c4_bench_$m.json 2> gpurun_out/c4_bench_$m.err
done
TMPI_GEMM_BULK=7 timeout 300 python bench.py --model alexnet --steps 20 --warmup 5 --repeats 5 > gpurun_out/c4_bench_alexnet_bulk7.json 2> gpurun_out/c4_bench_alexnet_bulk7.err
timeout 300 python bench.py --model alexnet --dtype tf32 --steps 20 --warmup 5 --repeats 5 > gpurun_out/c4_bench_alexnet_tf32.json 2> gpurun_out/c4_bench_alexnet_tf32.err
cat gpurun_out/c4_bench_*.json | cut -c1-330
for f in gpurun_out/c4_bench_*.err; do echo "== $f"; tail -5 $f | cut -c1-300; done
# per-launch device times of one eager step (cold cache, serialised: compare shares)
for m in resnet50 googlenet; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/c4_launches_$m.csv \
     python bench.py --model $m --steps 2 --warmup 3 --repeats 1 --no-graph > gpurun_out/c4_ncu_$m.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/c4_launches_alexnet_tf32.csv \
     python bench.py --model alexnet --dtype tf32 --steps 2 --warmup 3 --repeats 1 --no-graph > gpurun_out/c4_ncu_alexnet_tf32.log 2>&1
# full captures: the tf32 GEMM / conv kernel and the BN kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 60 -c 6 -o gpurun_out/c4_prof_gemm_tf32 \
     python bench.py --model alexnet --dtype tf32 --steps 2 --warmup 3 --repeats 1 --no-graph > gpurun_out/c4_ncu_full_tf32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bn_ -s 200 -c 6 -o gpurun_out/c4_prof_bn \
     python bench.py --model resnet50 --steps 2 --warmup 3 --repeats 1 --no-graph > gpurun_out/c4_ncu_full_bn.log 2>&1
ls -la gpurun_out/c4_*
