"""Multi-GPU tests (need >= 2 CUDA devices on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, script, *args, port=29611, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script)] + list(args)
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fused_kernels_two_ranks():
    r = _torchrun(2, "tests/mp_fused_check.py")
    assert r.returncode == 0 and "MP_FUSED_CHECK_OK" in r.stdout, r.stdout[-4000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("case", ["strategies", "easgd", "gosgd", "barrier"])
def test_protocols_and_strategies_numeric(case):
    """Per-strategy numeric checks on GPUs (the reference's test-exchanger suite), EASGD lost-update stress through the
    device-side ticket lock, GOSGD push-sum conservation under heavy gossip, 10,000-barrier stress with skewed launches."""
    n = min(4, torch.cuda.device_count()) if case in ("easgd", "gosgd") else 2
    r = _torchrun(n, "tests/mp_proto_check.py", case, port=29613, timeout=600)
    assert r.returncode == 0 and "MP_PROTO_CHECK_OK" in r.stdout, r.stdout[-4000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_bench_two_ranks_fused_matches_loss_scale():
    r = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "4", "--warmup", "3", port=29612)
    assert r.returncode == 0 and '"n_gpus": 2' in r.stdout, r.stdout[-4000:]


def _run_rule(rule_cls, devices, cfg=None, env=None, timeout=240):
    import subprocess
    rule = rule_cls()
    rule.model_config = dict(batch_size=64, file_batch_size=64, n_epochs=1, learning_rate=0.001, max_batches=12, printFreq=4,
                             data_kwargs=dict(n_synthetic=2048, synthetic=True))
    rule.model_config.update(cfg or {})
    rule.env.update(env or {})
    rule.init(devices=devices, modelfile="theanompi_b200.models.cifar10", modelclass="Cifar10_model")
    try:
        return rule.proc.wait(timeout=timeout)
    except subprocess.TimeoutExpired:
        rule.proc.kill()
        raise


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("strategy", ["fused", "fused_rs", "nccl32", "p2p32", "asa32", "asa16", "copper", "copper16", "nccl16", "ar"])
def test_rule_bsp_gpu(tmp_path, monkeypatch, strategy):
    import theanompi_b200 as tm
    monkeypatch.chdir(tmp_path)
    tm.BSP.sync_type, tm.BSP.exch_strategy = "cdd", strategy
    assert _run_rule(tm.BSP, ["cuda0", "cuda1"]) == 0
    tm.BSP.exch_strategy = "fused"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_rule_easgd_gpu(tmp_path, monkeypatch):
    import theanompi_b200 as tm
    monkeypatch.chdir(tmp_path)
    n = min(3, torch.cuda.device_count())
    assert _run_rule(tm.EASGD, ["cuda%d" % i for i in range(n)], env={"TMPI_EASGD_TAU": "4"}) == 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_rule_gosgd_gpu(tmp_path, monkeypatch):
    import theanompi_b200 as tm
    monkeypatch.chdir(tmp_path)
    n = min(4, torch.cuda.device_count())
    assert _run_rule(tm.GOSGD, ["cuda%d" % i for i in range(n)], cfg=dict(gosgd_p=0.3)) == 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fused_exchange_follows_classic_trajectory():
    """Training curves of the default fused exchange (owner-keeps-master, NVLS or P2P) and of the classic per-tensor NCCL
    strategy coincide (same data, same init): the reference's "1/2/4/8-GPU curves coincide" property, and a guard against
    any rank computing with stale parameters."""
    import json
    curves = {}
    for k, (name, strat, env) in enumerate((("classic", "nccl32", {}), ("fused", "fused", {}), ("fused_p2p", "fused", {"TMPI_NVLS": "0"}))):
        e = dict(os.environ); e.update(env)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(29640 + k), os.path.join(ROOT, "scripts/convergence.py"), "--steps", "100", "--bsp", "--strategy", strat]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, cwd=ROOT, env=e)
        line = [l for l in r.stdout.splitlines() if l.startswith("CONVERGENCE ")]
        assert r.returncode == 0 and line, r.stdout[-3000:]
        curves[name] = json.loads(line[-1][len("CONVERGENCE "):])["bsp_bf16"]["curve"]
    for name in ("fused", "fused_p2p"):
        for a, b in zip(curves["classic"], curves[name]):
            assert abs(a[1] - b[1]) < 0.08 + 0.1 * a[1], (name, curves)       # smoothed training loss
            assert abs(a[2] - b[2]) < 0.08 + 0.15 * a[2], (name, curves)      # validation cost
    assert curves["classic"][-1][1] < 0.2 * curves["classic"][0][1]             # and it actually learns
