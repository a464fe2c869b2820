#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "=== kernel tests (1 GPU)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=short --timeout 120 -p no:cacheprovider 2>&1 | tail -4
echo "=== multigpu tests"; timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q -x --tb=short --timeout 600 -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/t_multi2.log
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 "$@" 2>&1 | grep '^{' | tail -1 | cut -c1-260 | tee gpurun_out/bench2_$name.log; }
echo "=== bench N=2 default"; run default
echo "=== bench N=2 no-overlap"; run noov --no-overlap
echo "=== bench N=2 nccl32"; run nccl32 --strategy nccl32
