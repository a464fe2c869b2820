#!/bin/bash
# 2-GPU call b: owner-keeps-master exchange, EASGD / GOSGD benches, the other models' BSP rows at 2 GPUs
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29711 tests/mp_fused_check.py > gpurun_out/mg2b_fused_check.log 2>&1; tail -2 gpurun_out/mg2b_fused_check.log | cut -c1-600
run() { name=$1; shift; timeout 300 $TR --master-port $PORT bench.py --gpus 2 "$@" > gpurun_out/mg2b_$name.json 2> gpurun_out/mg2b_$name.err; PORT=$((PORT+1)); }
PORT=29720
run alexnet_okm --steps 20 --warmup 5 --repeats 5
TMPI_PUSH_MASTER=1 run alexnet_pushmaster --steps 20 --warmup 5 --repeats 3
run alexnet_fused16 --steps 20 --warmup 5 --repeats 3 --strategy fused16
TMPI_OVERLAP_BLOCKS=32 run alexnet_okm_b32 --steps 20 --warmup 5 --repeats 3
TMPI_OVERLAP_BLOCKS=96 run alexnet_okm_b96 --steps 20 --warmup 5 --repeats 3
run vgg16_easgd --steps 16 --warmup 4 --repeats 3 --model vgg16 --rule easgd --tau 4
run wrn_gosgd --steps 20 --warmup 5 --repeats 3 --model wrn --rule gosgd
run googlenet --steps 20 --warmup 5 --repeats 3 --model googlenet
run vgg16 --steps 20 --warmup 5 --repeats 3 --model vgg16
run resnet50 --steps 20 --warmup 5 --repeats 3 --model resnet50
for f in gpurun_out/mg2b_*.json; do echo "== $f"; cut -c1-420 $f; done
for f in gpurun_out/mg2b_*.err; do echo "== $f"; grep -v "OMP_NUM\|^\*\*\*\|^$" $f | tail -6 | cut -c1-300; done
