#!/bin/bash
# what the driver runs at round end on one GPU: all gpu-marked tests, smoke(), bench.py (both arms)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/t_all_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3
echo "=== bench reference arm"; timeout 120 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-300
echo "=== bench"; timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench1_final.json | cut -c1-1200
