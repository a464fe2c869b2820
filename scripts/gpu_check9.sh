#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== model tests"; timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/t_models.log
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches5.csv python scripts/profile_step.py > gpurun_out/ncu_launch5.log 2>&1; tail -2 gpurun_out/ncu_launch5.log
python scripts/summarize_launches.py gpurun_out/launches5.csv gpurun_out/step_order5.txt | head -30
