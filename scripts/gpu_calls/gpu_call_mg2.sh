#!/bin/bash
# 2-GPU call: multi-GPU test-suite (fused kernels, protocol stress, per-strategy numerics, rules) + 2-GPU benches + NVLink counters
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/mg2_topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_models.py -q -k "tf32 or loader_process" > gpurun_out/mg2_pytest_tf32.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/mg2_pytest_tf32.log | tail -12
timeout 1500 python -m pytest tests/test_multigpu.py -q > gpurun_out/mg2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/mg2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi nvlink -gt d > gpurun_out/mg2_nvl_before_fused.txt 2>&1
timeout 300 $TR --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 5 > gpurun_out/mg2_bench_alexnet_fused.json 2> gpurun_out/mg2_bench_alexnet_fused.err
nvidia-smi nvlink -gt d > gpurun_out/mg2_nvl_after_fused.txt 2>&1
timeout 300 $TR --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 5 --strategy fused_rs > gpurun_out/mg2_bench_alexnet_fused_rs.json 2> gpurun_out/mg2_bench_alexnet_fused_rs.err
nvidia-smi nvlink -gt d > gpurun_out/mg2_nvl_after_rs.txt 2>&1
timeout 300 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 --no-overlap > gpurun_out/mg2_bench_alexnet_noov.json 2> gpurun_out/mg2_bench_alexnet_noov.err
timeout 300 $TR --master-port 29704 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 --impl torch_best > gpurun_out/mg2_torch_alexnet.json 2> gpurun_out/mg2_torch_alexnet.err
timeout 300 $TR --master-port 29705 bench.py --gpus 2 --steps 16 --warmup 4 --repeats 3 --model vgg16 --rule easgd --tau 4 > gpurun_out/mg2_bench_vgg16_easgd.json 2> gpurun_out/mg2_bench_vgg16_easgd.err
timeout 300 $TR --master-port 29706 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 --model wrn --rule gosgd > gpurun_out/mg2_bench_wrn_gosgd.json 2> gpurun_out/mg2_bench_wrn_gosgd.err
timeout 300 $TR --master-port 29707 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 --dtype tf32 > gpurun_out/mg2_bench_alexnet_tf32.json 2> gpurun_out/mg2_bench_alexnet_tf32.err
grep -E "passed|failed|FAILED|rc=" gpurun_out/mg2_pytest.log | tail -20
cat gpurun_out/mg2_bench_*.json gpurun_out/mg2_torch_*.json | cut -c1-360
for f in gpurun_out/mg2_*.err; do echo "== $f"; tail -6 $f | cut -c1-300; done
