#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== conv tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -p no:cacheprovider -k "conv" 2>&1 | tail -40 | tee gpurun_out/t_conv.log
echo "=== all kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -p no:cacheprovider -k "not conv" 2>&1 | tail -6
echo "=== bench N=1 implicit"; timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -2 | cut -c1-400 | tee gpurun_out/bench1_implicit.log
echo "=== bench N=1 explicit"; TMPI_CONV=explicit timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/bench1_explicit.log
echo "=== model tests"; timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/t_models.log
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches4.csv python scripts/profile_step.py > gpurun_out/ncu_launch4.log 2>&1; tail -2 gpurun_out/ncu_launch4.log
python scripts/summarize_launches.py gpurun_out/launches4.csv gpurun_out/step_order4.txt | head -24
