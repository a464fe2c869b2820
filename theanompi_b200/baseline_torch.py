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


# =========================================================================== the strongest same-semantics LIBRARY build
class TorchVGG16(nn.Module):
    def __init__(self, n_class=1000):
        super().__init__()
        cfg = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
        layers, cin = [], 3
        for v in cfg:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, 1, 1), nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*layers)
        self.cls = nn.Sequential(nn.Linear(512 * 7 * 7, 4096), nn.ReLU(inplace=True), nn.Dropout(0.5),
                                 nn.Linear(4096, 4096), nn.ReLU(inplace=True), nn.Dropout(0.5), nn.Linear(4096, n_class))

    def forward(self, x):
        return self.cls(self.features(x).flatten(1))


class _Incept(nn.Module):
    def __init__(self, cin, n1, r3, n3, r5, n5, pj):
        super().__init__()
        self.b1 = nn.Conv2d(cin, n1, 1)
        self.b3 = nn.Sequential(nn.Conv2d(cin, r3, 1), nn.ReLU(inplace=True), nn.Conv2d(r3, n3, 3, 1, 1))
        self.b5 = nn.Sequential(nn.Conv2d(cin, r5, 1), nn.ReLU(inplace=True), nn.Conv2d(r5, n5, 5, 1, 2))
        self.bp = nn.Sequential(nn.MaxPool2d(3, 1, 1), nn.Conv2d(cin, pj, 1))

    def forward(self, x):
        return F.relu(torch.cat([self.b1(x), self.b3(x), self.b5(x), self.bp(x)], 1))


class _Aux(nn.Module):
    def __init__(self, cin, n_class):
        super().__init__()
        self.conv = nn.Conv2d(cin, 128, 1)
        self.f1, self.f2 = nn.Linear(128 * 4 * 4, 1024), nn.Linear(1024, n_class)

    def forward(self, x):
        x = F.relu(self.conv(F.avg_pool2d(x, 5, 3)))
        return self.f2(F.dropout(F.relu(self.f1(x.flatten(1))), 0.7))


class TorchGoogLeNet(nn.Module):
    """GoogLeNet with both auxiliary towers (ref ``models/googlenet.py:46-181``); returns main + 0.3·aux losses' logits."""

    def __init__(self, n_class=1000):
        super().__init__()
        self.c1 = nn.Conv2d(3, 64, 7, 2, 3)
        self.c2r, self.c2 = nn.Conv2d(64, 64, 1), nn.Conv2d(64, 192, 3, 1, 1)
        self.i3a, self.i3b = _Incept(192, 64, 96, 128, 16, 32, 32), _Incept(256, 128, 128, 192, 32, 96, 64)
        self.i4a = _Incept(480, 192, 96, 208, 16, 48, 64)
        self.i4b = _Incept(512, 160, 112, 224, 24, 64, 64)
        self.i4c = _Incept(512, 128, 128, 256, 24, 64, 64)
        self.i4d = _Incept(512, 112, 144, 288, 32, 64, 64)
        self.i4e = _Incept(528, 256, 160, 320, 32, 128, 128)
        self.i5a, self.i5b = _Incept(832, 256, 160, 320, 32, 128, 128), _Incept(832, 384, 192, 384, 48, 128, 128)
        self.aux1, self.aux2 = _Aux(512, n_class), _Aux(528, n_class)
        self.fc = nn.Linear(1024, n_class)

    def forward(self, x):
        x = F.local_response_norm(F.max_pool2d(F.relu(self.c1(x)), 3, 2, ceil_mode=True), 5, 1e-4 * 5, 0.75, 2.0)
        x = F.relu(self.c2(F.relu(self.c2r(x))))
        x = F.max_pool2d(F.local_response_norm(x, 5, 1e-4 * 5, 0.75, 2.0), 3, 2, ceil_mode=True)
        x = F.max_pool2d(self.i3b(self.i3a(x)), 3, 2, ceil_mode=True)
        x = self.i4a(x); a1 = self.aux1(x)
        x = self.i4d(self.i4c(self.i4b(x))); a2 = self.aux2(x)
        x = F.max_pool2d(self.i4e(x), 3, 2, ceil_mode=True)
        x = self.i5b(self.i5a(x))
        x = F.dropout(F.adaptive_avg_pool2d(x, 1).flatten(1), 0.4)
        return self.fc(x), a1, a2


def _torch_model(name):
    if name == "alexnet":
        return TorchAlexNet(), (227, 227), 1000
    if name == "vgg16":
        return TorchVGG16(), (224, 224), 1000
    if name == "googlenet":
        return TorchGoogLeNet(), (224, 224), 1000
    if name == "resnet50":
        from .models.lasagne_model_zoo.resnet50 import ResNet50Net
        return ResNet50Net(), (224, 224), 1000
    if name == "wrn":
        from .models.keras_model_zoo.wresnet import WRN
        return WRN(28, 4, 10), (32, 32), 10
    raise ValueError(name)


def run_torch_best(args, rank, world, local, K, Wm, ClockSampler, published):
    """cuDNN / cuBLAS channels-last (bf16 autocast, or fp32 storage with TF32 math for ``--dtype tf32``), the WHOLE step
    (forward, backward, one flat-bucket ``ncclAllReduce`` of the gradient arena, a 3-kernel flat momentum-SGD update) captured
    in ONE CUDA graph.  Same update rule as the product (``u = μu + g/k + wd·w; w −= lr·u``)."""
    import statistics
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(23455)
    tf32 = args.dtype == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    net, hw, n_class = _torch_model(args.model)
    net = net.to(dev).to(memory_format=torch.channels_last)
    params = [p for p in net.parameters() if p.requires_grad]
    n = sum(p.numel() for p in params)
    Wf, Gf, Uf = (torch.zeros(n, device=dev) for _ in range(3))
    wdmask = torch.zeros(n, device=dev)
    off = 0
    with torch.no_grad():
        for p in params:
            k = p.numel()
            Wf[off:off + k].copy_(p.reshape(-1))
            # contiguous flat storage for weights and grads (channels-last conv weights keep their strides inside the view)
            p.data = Wf[off:off + k].view(p.shape) if p.dim() != 4 else p.data
            if p.dim() == 4:
                v = Wf[off:off + k].view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
                v.copy_(p.data); p.data = v
                p.grad = Gf[off:off + k].view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
            else:
                p.grad = Gf[off:off + k].view(p.shape)
            wdmask[off:off + k] = 5e-4 if p.dim() > 1 else 0.0
            off += k
    lr, mu = 0.01, 0.9
    B = args.batch
    x = torch.randn(B, 3, hw[0], hw[1], device=dev).to(memory_format=torch.channels_last)
    y = torch.randint(0, n_class, (B,), device=dev)
    raw_hw = 256 if hw[0] > 32 else 32
    pinned = torch.empty((B, raw_hw, raw_hw, 3), dtype=torch.uint8).pin_memory()
    stage = torch.empty((B, raw_hw, raw_hw, 3), dtype=torch.uint8, device=dev)
    o = (raw_hw - hw[0]) // 2
    loss_buf = torch.zeros((), device=dev)

    def body():
        Gf.zero_()
        if tf32:
            out = net(x)
        else:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(x)
        if isinstance(out, tuple):
            loss = F.cross_entropy(out[0].float(), y) + 0.3 * F.cross_entropy(out[1].float(), y) + 0.3 * F.cross_entropy(out[2].float(), y)
        else:
            loss = F.cross_entropy(out.float(), y)
        loss.backward()
        with torch.no_grad():
            if world > 1:
                dist.all_reduce(Gf)
            Uf.mul_(mu).add_(Gf, alpha=1.0 / world).addcmul_(wdmask, Wf)
            Wf.sub_(Uf, alpha=lr)
            loss_buf.copy_(loss.detach())

    def e2e_stage():
        stage.copy_(pinned, non_blocking=True)
        x.copy_(stage[:, o:o + hw[0], o:o + hw[1], :].permute(0, 3, 1, 2).float().div_(255))

    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = None
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            body()
        graph = g
    except Exception as e:  # noqa: BLE001
        print("[torch_best] CUDA graph capture failed (%s) — eager" % (repr(e)[:200],))
        torch.cuda.synchronize()
    step = (lambda: graph.replay()) if graph is not None else body

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):
        step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    R = max(1, getattr(args, "repeats", 5))
    dev_ms, e2e_ms = [], []
    for _ in range(R):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); e0.record()
        for _ in range(K):
            step()
        e1.record(); barrier()
        dev_ms.append(e0.elapsed_time(e1) / K)
    for _ in range(Wm):
        e2e_stage(); step()
    lv = 0.0
    for _ in range(R):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); e0.record()
        for _ in range(K):
            e2e_stage(); step()
            lv = float(loss_buf)
        e1.record(); barrier()
        e2e_ms.append(e0.elapsed_time(e1) / K)
    clocks = sampler.stop()
    t = torch.tensor(dev_ms + e2e_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = [float(v) for v in t[:R]], [float(v) for v in t[R:]]
    ms, ms2 = statistics.median(dev_ms), statistics.median(e2e_ms)
    per = 5120.0 / (B * world)
    if rank == 0:
        base = published.get(world)
        print(json.dumps({
            "metric": "%s-%db BSP seconds per 5120 images (train+comm, device-timed, max over ranks)" % (args.model, B),
            "value": ms * per / 1000.0, "unit": "s/5120img", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": (ms * per / 1000.0 / base) if base else None, "dtype": args.dtype, "data": "synthetic",
            "impl": "torch_best (cuDNN/cuBLAS channels-last + flat ncclAllReduce + flat momentum-SGD, CUDA graph: %s)" % (graph is not None),
            "repeats": {"n": R, "ms_per_step_min": min(dev_ms), "ms_per_step_max": max(dev_ms)},
            "config": {"model": args.model, "global_batch": B * world, "parallelism": "dp%d" % world},
            "clocks": clocks,
            "e2e": {"value": ms2 * per / 1000.0, "unit": "s/5120img", "ms_per_step": ms2, "h2d_bytes_per_step": pinned.numel(),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": 0, "final_loss": lv}), flush=True)
    if world > 1:
        # a process group whose collectives were captured into a CUDA graph can hang in destroy_process_group(): leave at once
        import os
        import sys
        barrier()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
    return 0
