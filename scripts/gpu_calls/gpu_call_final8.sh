#!/bin/bash
# last 8-GPU call of the round: 8-GPU training curve (same global batch as the 1- and 2-GPU runs) + ResNet50 with the final BN kernels
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29990 scripts/convergence.py --steps 100 --bsp > gpurun_out/final_convergence8.log 2>&1; grep CONVERGENCE gpurun_out/final_convergence8.log | cut -c1-500
timeout 120 $TR --master-port 29991 bench.py --gpus $N --model resnet50 --steps 10 --warmup 3 --repeats 3 > gpurun_out/final_resnet50_8gpu.json 2> gpurun_out/final_resnet50_8gpu.err; tail -c 2500 gpurun_out/final_resnet50_8gpu.json | cut -c1-400
grep -i "error\|Traceback" gpurun_out/final_*.err gpurun_out/final_convergence8.log | head -5
