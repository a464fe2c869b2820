#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
echo "=== kernel tests (1 GPU)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/t_kernels.log
echo "=== bench N=1"; timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench1.log
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches2.csv python scripts/profile_step.py > gpurun_out/ncu_launch2.log 2>&1; tail -2 gpurun_out/ncu_launch2.log
python scripts/summarize_launches.py gpurun_out/launches2.csv gpurun_out/step_order2.txt | head -30
echo "=== 2-GPU fused kernel check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mp_fused_check.py --sweep 2>&1 | tail -25 | tee gpurun_out/mp_fused.log
echo "=== bench N=2 fused (overlap)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 40 --warmup 5 2>&1 | tail -4 | tee gpurun_out/bench2_fused.log
echo "=== bench N=2 fused no overlap"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 40 --warmup 5 --no-overlap 2>&1 | tail -2 | tee gpurun_out/bench2_noov.log
echo "=== bench N=2 nccl32 strategy (our compute)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 40 --warmup 5 --strategy nccl32 2>&1 | tail -2 | tee gpurun_out/bench2_nccl32.log
echo "=== bench N=2 torch/nccl baseline"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 2 --steps 40 --warmup 5 --impl nccl_baseline 2>&1 | tail -2 | tee gpurun_out/bench2_base.log
