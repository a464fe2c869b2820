#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29460
go() { tag=$1; shift; timeout 120 $TR --master-port $P scripts/convergence.py --steps 100 --bsp "$@" > gpurun_out/fix2_$tag.log 2>&1; echo "$tag: $(grep CONVERGENCE gpurun_out/fix2_$tag.log | cut -c40-400)"; grep -i "error" gpurun_out/fix2_$tag.log | head -3; P=$((P+1)); }
go classic --strategy nccl32
go fused
TMPI_NVLS=0 go fused_p2p
TMPI_PUSH_MASTER=1 go fused_pm1
TMPI_ONESHOT_BYTES=0 go fused_twoshot
timeout 200 $TR --master-port 29470 tests/mp_fused_check.py > gpurun_out/fix2_mpfused.log 2>&1; tail -2 gpurun_out/fix2_mpfused.log
