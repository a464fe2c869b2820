#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29420
go() { tag=$1; shift; timeout 100 $TR --master-port $P scripts/convergence.py --steps 60 --bsp "$@" > gpurun_out/dbg2_$tag.log 2>&1; echo "$tag: $(grep CONVERGENCE gpurun_out/dbg2_$tag.log | cut -c60-330)"; P=$((P+1)); }
go A_default
go B_default
TMPI_PUSH_MASTER=1 go C_pm1
go D_pm0_noov --no-overlap
TMPI_PUSH_MASTER=1 TMPI_ONESHOT_BYTES=0 go E_pm1_twoshot
TMPI_ONESHOT_BYTES=0 go F_pm0_twoshot
TMPI_PUSH_MASTER=1 go G_pm1_noov --no-overlap
