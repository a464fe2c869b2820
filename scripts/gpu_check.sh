#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=short --timeout 120 -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/t_kernels.log
echo "=== model tests"; timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/t_models.log
echo "=== bench N=1"; timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/bench1_elt.log
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches6.csv python scripts/profile_step.py > gpurun_out/ncu_launch6.log 2>&1; tail -2 gpurun_out/ncu_launch6.log
python scripts/summarize_launches.py gpurun_out/launches6.csv gpurun_out/step_order6.txt | head -24
