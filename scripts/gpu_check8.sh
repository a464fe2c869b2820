#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== gemm/conv tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -p no:cacheprovider -k "conv or gemm or linear" 2>&1 | tail -8 | tee gpurun_out/t_conv.log
echo "=== probe"; timeout 600 python scripts/gemm_probe.py 2>&1 | tail -30 | tee gpurun_out/gemm_probe.txt
echo "=== bench N=1"; timeout 300 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-500 | tee gpurun_out/bench1_split.log
