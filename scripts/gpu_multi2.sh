#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== easgd/gosgd rule tests"; timeout 700 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -k "easgd or gosgd" 2>&1 | tail -40 | tee gpurun_out/t_multi2b.log
run() { name=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 "$@" 2>&1 | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config'].get('overlap'), d['config'].get('exch_strategy'))" | tee gpurun_out/bench2_$name.log; }
echo "=== N=2 no-overlap"; run noov --no-overlap
echo "=== N=2 overlap blocks 64"; TMPI_OVERLAP_BLOCKS=64 run ov64
echo "=== N=2 overlap blocks 148"; TMPI_OVERLAP_BLOCKS=148 run ov148
echo "=== N=2 overlap blocks 16"; TMPI_OVERLAP_BLOCKS=16 run ov16
echo "=== N=2 fused16 no-overlap"; run f16 --no-overlap --strategy fused16
