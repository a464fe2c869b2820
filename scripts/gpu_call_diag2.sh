#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29400
for v in "fused 0 1" "fused 1 1" "fused 0 0" "nccl32 0 1"; do
  set -- $v
  TMPI_PUSH_MASTER=$2 TMPI_NVLS=$3 timeout 120 $TR --master-port $P scripts/convergence.py --steps 120 --bsp --strategy $1 > gpurun_out/diag2_$1_$2_$3.log 2>&1
  grep CONVERGENCE gpurun_out/diag2_$1_$2_$3.log | cut -c1-500
  P=$((P+1))
done
