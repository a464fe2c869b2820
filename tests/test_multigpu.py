"""Multi-GPU tests (need >= 2 CUDA devices on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, script, *args, port=29611, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script)] + list(args)
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fused_kernels_two_ranks():
    r = _torchrun(2, "tests/mp_fused_check.py")
    assert r.returncode == 0 and "MP_FUSED_CHECK_OK" in r.stdout, r.stdout[-4000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_bench_two_ranks_fused_matches_loss_scale():
    r = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "4", "--warmup", "3", port=29612)
    assert r.returncode == 0 and '"n_gpus": 2' in r.stdout, r.stdout[-4000:]
