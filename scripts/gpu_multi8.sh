#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "=== fused check x8"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/mp_fused_check.py 2>&1 | tail -6 | tee gpurun_out/mp_fused_8.log
run() { n=$1; name=$2; shift; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 40 --warmup 5 "$@" 2>&1 | grep '^{' | tail -1 > gpurun_out/bench${n}_$name.json; python -c "import sys,json; d=json.loads(open('gpurun_out/bench${n}_$name.json').read()); print('$n $name', d['ms_per_step'], d['value'], d['config'].get('overlap'), d['config'].get('exch_strategy'), d.get('clocks'))"; }
echo "=== N=8 default"; run 8 default
echo "=== N=8 no-overlap"; run 8 noov --no-overlap
echo "=== N=8 overlap 96"; TMPI_OVERLAP_BLOCKS=96 run 8 ov96
echo "=== N=4 default"; run 4 default
echo "=== N=8 nccl32"; run 8 nccl32 --strategy nccl32
