#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_multigpu.py -q -k "follows_classic or fused_kernels" > gpurun_out/diag2b_pytest.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/diag2b_pytest.log | tail -8
timeout 240 $TR --master-port 29770 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 5 > gpurun_out/diag2b_alexnet.json 2> gpurun_out/diag2b_alexnet.err; cut -c1-300 gpurun_out/diag2b_alexnet.json
timeout 240 $TR --master-port 29771 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 --model resnet50 > gpurun_out/diag2b_resnet50.json 2> gpurun_out/diag2b_resnet50.err; cut -c1-300 gpurun_out/diag2b_resnet50.json
