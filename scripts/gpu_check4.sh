#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mp_fused_check.py > gpurun_out/mp_fused_full.log 2>&1
grep -n -B2 -A25 "Traceback" gpurun_out/mp_fused_full.log | head -120
tail -5 gpurun_out/mp_fused_full.log
