"""Micro-benchmark of one model's training function (ref ``theanompi/models/be_model.py:62-77``):
100 × ``train_iter_fn(0)`` with a device sync on both sides, CUDA-event timed.

    python -m theanompi_b200.models.be_model <modelfile> <modelclass> [n_iters]
"""
from __future__ import annotations

import sys
import time

import torch


def be_model(modelfile, modelclass, n_iters=100, config=None):
    import importlib
    cfg = dict(verbose=False, rank=0, size=1, mname=modelclass, cuda_graph=True)
    cfg.update(config or {})
    model = getattr(importlib.import_module(modelfile), modelclass)(cfg)
    model.compile_iter_fns(sync_type="avg")
    cuda = model.device.type == "cuda"
    for _ in range(5):
        model.train_iter_fn(0)
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.time()
    for _ in range(n_iters):
        model.train_iter_fn(0)
    if cuda:
        e1.record(); torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1000.0
    else:
        sec = time.time() - t0
    print("%s: %d iterations in %.4f s (%.3f ms/iter, %.1f images/s)" % (modelclass, n_iters, sec, 1000 * sec / n_iters,
                                                                       n_iters * model.batch_size / sec))
    model.cleanup()
    return sec / n_iters


if __name__ == "__main__":
    a = sys.argv[1:]
    be_model(a[0], a[1], int(a[2]) if len(a) > 2 else 100)
