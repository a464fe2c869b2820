#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=short --timeout 120 -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/t_kernels.log
echo "=== bench N=1"; timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench1_fc.log | cut -c1-330
echo "=== bench N=1 no fc splitk"; TMPI_FC_SPLITK=0 timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-330
