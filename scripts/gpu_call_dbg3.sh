#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29440
go() { tag=$1; shift; timeout 100 $TR --master-port $P scripts/convergence.py --steps 60 --bsp "$@" > gpurun_out/dbg3_$tag.log 2>&1; echo "$tag: $(grep CONVERGENCE gpurun_out/dbg3_$tag.log | cut -c60-300)"; grep -i "error\|capture" gpurun_out/dbg3_$tag.log | head -3; P=$((P+1)); }
export TMPI_PUSH_MASTER=1
go H_ov_blocks296 --comm-blocks 296
CUDA_LAUNCH_BLOCKING=1 go I_ov_blocking
go J_graph_ov --graph
go K_graph_noov --graph --no-overlap
