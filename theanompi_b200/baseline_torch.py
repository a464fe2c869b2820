"""Reference-*semantics* baseline inside this framework (BASELINE.md "How this maps").

The reference cannot run in this image, so the live yardstick reproduces what it does,
with today's libraries on the same box: library compute (torch → cuDNN / cuBLAS, bf16
autocast, channels-last), then — strictly after backward — ``Barrier`` → one
``ncclAllReduce`` per parameter tensor (``exchanger_strategy.py:121-127``) → separate
elementwise update kernels per tensor (``opt.py:181-268``).  No fused kernels, no overlap,
no CUDA graph.  ``bench.py --impl nccl_baseline`` prints the same JSON line as the product.
"""
from __future__ import annotations

import json

import torch
import torch.nn as nn
import torch.nn.functional as F


class TorchAlexNet(nn.Module):
    def __init__(self, n_class=1000):
        super().__init__()
        self.c1 = nn.Conv2d(3, 96, 11, 4)
        self.c2 = nn.Conv2d(96, 256, 5, 1, 2, groups=2)
        self.c3 = nn.Conv2d(256, 384, 3, 1, 1)
        self.c4 = nn.Conv2d(384, 384, 3, 1, 1, groups=2)
        self.c5 = nn.Conv2d(384, 256, 3, 1, 1, groups=2)
        self.f6, self.f7, self.f8 = nn.Linear(9216, 4096), nn.Linear(4096, 4096), nn.Linear(4096, n_class)

    def forward(self, x):
        x = F.local_response_norm(F.max_pool2d(F.relu(self.c1(x)), 3, 2), 5, 1e-4 * 5, 0.75, 2.0)
        x = F.local_response_norm(F.max_pool2d(F.relu(self.c2(x)), 3, 2), 5, 1e-4 * 5, 0.75, 2.0)
        x = F.relu(self.c3(x)); x = F.relu(self.c4(x))
        x = F.max_pool2d(F.relu(self.c5(x)), 3, 2)
        x = x.flatten(1)
        x = F.dropout(F.relu(self.f6(x)), 0.5)
        x = F.dropout(F.relu(self.f7(x)), 0.5)
        return self.f8(x)


def run_baseline(args, rank, world, local, K, Wm, ClockSampler, published):
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(23455)
    net = TorchAlexNet().to(dev).to(memory_format=torch.channels_last)
    params = list(net.parameters())
    vels = [torch.zeros_like(p) for p in params]
    vels2 = [torch.zeros_like(p) for p in params]
    lr, mu, wd = 0.01, 0.9, 5e-4
    B = args.batch
    x = torch.randn(B, 3, 227, 227, device=dev).to(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (B,), device=dev)
    pinned = torch.empty((B, 256, 256, 3), dtype=torch.uint8).pin_memory()

    def step(e2e=False):
        if e2e:                                             # reference loader: H2D of the batch, crop on device
            raw = pinned.to(dev, non_blocking=True)
            xb = raw[:, 14:241, 14:241, :].permute(0, 3, 1, 2).float().div_(255).contiguous(memory_format=torch.channels_last)
        else:
            xb = x
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(net(xb).float(), y)
        for p in params:
            p.grad = None
        loss.backward()
        with torch.no_grad():
            for p, u in zip(params, vels):                  # pre: u = mu*u + (g + wd*w)
                g = p.grad + wd * p if p.dim() > 1 else p.grad
                u.mul_(mu).add_(g)
            if world > 1:
                dist.barrier()
                for u, r in zip(vels, vels2):               # one NCCL call per tensor
                    r.copy_(u)
                    dist.all_reduce(r)
                for p, r in zip(params, vels2):             # post: w -= lr * r / k
                    p.sub_(r, alpha=(lr if p.dim() > 1 else 2 * lr) / world)
            else:
                for p, u in zip(params, vels):
                    p.sub_(u, alpha=lr if p.dim() > 1 else 2 * lr)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):
        step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); e0.record()
    for _ in range(K):
        loss = step()
    e1.record(); barrier()
    ms = e0.elapsed_time(e1)
    for _ in range(Wm):
        step(True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(K):
        lv = float(step(True))
    f1.record(); barrier()
    ms2 = f0.elapsed_time(f1)
    clocks = sampler.stop()
    t = torch.tensor([ms, ms2], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms2 = float(t[0]), float(t[1])
    per = 5120.0 / (B * world)
    if rank == 0:
        base = published.get(world)
        print(json.dumps({
            "metric": "AlexNet-128b BSP seconds per 5120 images (train+comm, device-timed, max over ranks)",
            "value": ms / K * per / 1000.0, "unit": "s/5120img", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": (ms / K * per / 1000.0 / base) if base else None, "dtype": "bf16", "data": "synthetic",
            "impl": "nccl_baseline (torch cuDNN/cuBLAS + per-tensor ncclAllReduce + per-tensor updates)",
            "config": {"model": "AlexNet", "global_batch": B * world, "parallelism": "dp%d" % world},
            "clocks": clocks,
            "e2e": {"value": ms2 / K * per / 1000.0, "unit": "s/5120img", "h2d_bytes_per_step": pinned.numel(),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": 0, "final_loss": lv}))
    if world > 1:
        dist.destroy_process_group()
    return 0
