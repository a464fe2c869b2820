#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29410 scripts/debug_okm.py > gpurun_out/dbg_okm.log 2>&1
grep -E "param |push_master|Error|error" gpurun_out/dbg_okm.log | cut -c1-250
