#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=short --timeout 120 -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/t_kernels.log
echo "=== bench N=1"; timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench1_wave.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['final_loss'], d['native_launches_per_step'])"
