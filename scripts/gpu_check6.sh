#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/t_kernels.log
echo "=== bench N=1"; timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench1_persist.log
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches3.csv python scripts/profile_step.py > gpurun_out/ncu_launch3.log 2>&1; tail -2 gpurun_out/ncu_launch3.log
python scripts/summarize_launches.py gpurun_out/launches3.csv gpurun_out/step_order3.txt | head -24
