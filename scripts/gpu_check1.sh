#!/bin/bash
# First GPU pass: kernel numerics, smoke, short bench.  Everything under its own timeout.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python - <<'PY' > gpurun_out/env.txt 2>&1
import torch; print(torch.__version__, torch.cuda.is_available(), torch.cuda.get_device_name(0))
from theanompi_b200.ops import native; print("native:", native.available(), native.load_error())
PY
echo "=== gemm tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm" --timeout 120 -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/t_gemm.log
echo "=== other kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not gemm" --timeout 120 -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/t_rest.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "=== bench eager"; timeout 600 python bench.py --steps 10 --warmup 3 --no-graph 2>&1 | tail -8 | tee gpurun_out/bench_eager.log
echo "=== bench graph"; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -8 | tee gpurun_out/bench_graph.log
echo "=== baseline"; timeout 600 python bench.py --impl nccl_baseline --steps 20 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_base.log
