#!/bin/bash
# 8-GPU call: every BASELINE.json config at 8 GPUs + the published table's 8-GPU column + protocol stress at 8 ranks
set -x
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29800
run() { name=$1; shift; timeout 240 $TR --master-port $PORT bench.py --gpus $N "$@" > gpurun_out/mg${N}_$name.json 2> gpurun_out/mg${N}_$name.err; PORT=$((PORT+1)); tail -c 600 gpurun_out/mg${N}_$name.json | cut -c1-300; }
nvidia-smi nvlink -gt d > gpurun_out/mg${N}_nvl_0.txt 2>&1
run alexnet --steps 20 --warmup 5 --repeats 5
nvidia-smi nvlink -gt d > gpurun_out/mg${N}_nvl_1.txt 2>&1
TMPI_FUSED_U=4 TMPI_PUSH_MASTER=1 run alexnet_r1cfg --steps 20 --warmup 5 --repeats 3
run googlenet --steps 20 --warmup 5 --repeats 3 --model googlenet
run vgg16 --steps 20 --warmup 5 --repeats 3 --model vgg16
run resnet50 --steps 20 --warmup 5 --repeats 3 --model resnet50
run vgg16_easgd --steps 16 --warmup 4 --repeats 3 --model vgg16 --rule easgd --tau 4
TMPI_EASGD_LOCKFREE=1 run vgg16_easgd_lockfree --steps 16 --warmup 4 --repeats 3 --model vgg16 --rule easgd --tau 4
run wrn_gosgd --steps 20 --warmup 5 --repeats 3 --model wrn --rule gosgd
run alexnet_tf32 --steps 20 --warmup 5 --repeats 3 --dtype tf32
run torch_alexnet --steps 20 --warmup 5 --repeats 3 --impl torch_best
timeout 300 $TR --master-port 29850 tests/mp_proto_check.py easgd gosgd > gpurun_out/mg${N}_proto.log 2>&1; tail -1 gpurun_out/mg${N}_proto.log | cut -c1-700
timeout 200 $TR --master-port 29990 scripts/convergence.py --steps 320 --bsp > gpurun_out/mg${N}_convergence.log 2>&1; grep CONVERGENCE gpurun_out/mg${N}_convergence.log | cut -c1-600
for f in gpurun_out/mg${N}_*.err; do echo "== $f"; grep -v "OMP_NUM\|^\*\*\*\|^$" $f | grep -i "error\|Traceback" | head -4; done
