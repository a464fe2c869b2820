#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=short --timeout 120 -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/t_kernels.log
echo "=== probe"; timeout 600 python scripts/gemm_probe.py --quick 2>&1 | tail -20 | tee gpurun_out/gemm_probe.txt
echo "=== bench N=1"; timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/bench1_tall.log
echo "=== bench N=1 no tall"; TMPI_GEMM_TALL=0 timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/bench1_notall.log
