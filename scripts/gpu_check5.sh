#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mp_fused_check.py --sweep > gpurun_out/mp_fused_full_$N.log 2>&1
grep -n -A25 "Traceback" gpurun_out/mp_fused_full_$N.log | head -60
grep -E "symmetric|MP_FUSED" gpurun_out/mp_fused_full_$N.log | cut -c1-1500
IFS=";" read -ra CFGS <<< "${BENCH_CFGS:-fused;fused --no-overlap;nccl32}"; for cfg in "${CFGS[@]}"; do
  set -- $cfg; s=$1; shift
  echo "=== bench N=$N strategy=$s $*"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 40 --warmup 5 --strategy $s $* > gpurun_out/bench${N}_${s}_$#.log 2>&1
  grep -E '^\{"metric"' gpurun_out/bench${N}_${s}_$#.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','images_per_s')}, d['config']['exch_strategy'], 'overlap', d['config']['overlap'], 'e2e', d['e2e']['ms_per_step'])" || tail -20 gpurun_out/bench${N}_${s}_$#.log
  grep -n -A12 "Traceback" gpurun_out/bench${N}_${s}_$#.log | head -40
done
