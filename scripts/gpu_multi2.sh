#!/bin/bash
# 2 GPUs: every multi-GPU test (fused-kernel check, BSP strategies, EASGD, GOSGD through the Rule API) + bench.py variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== multigpu tests"; timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/t_multi2.log
run() { name=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 "$@" 2>&1 | grep '^{' | tail -1 > gpurun_out/bench2_$name.json; python -c "import json; d=json.loads(open('gpurun_out/bench2_$name.json').read()); print('$name', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config'].get('overlap'), d['config'].get('exch_strategy'))"; }
echo "=== N=2 default"; run default
echo "=== N=2 no-overlap"; run noov --no-overlap
