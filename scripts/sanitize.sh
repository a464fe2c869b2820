#!/bin/bash
# compute-sanitizer over the hand-written kernels on tiny shapes (the reference has no race detection).
#   scripts/sanitize.sh [memcheck|racecheck|synccheck]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
tool=${1:-memcheck}
mkdir -p gpurun_out
timeout 240 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 120 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider \
  -k "lrn or test_pool or conv_pool_fused or dropout or softmax or crop or sgd_flat or legacy or test_gemm_epilogue or small_batch" 2>&1 | tail -25 | tee gpurun_out/sanitize_$tool.log
