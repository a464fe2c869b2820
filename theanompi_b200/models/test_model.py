"""Single-GPU model tester (ref ``theanompi/models/test_model.py``): "test your single GPU
model before trying a rule" (``README.md:126``) — the full train/val loop with a Recorder,
no process group, ``sync_type='avg'``.

    python -m theanompi_b200.models.test_model <modelfile> <modelclass> [device] [max_batches]
"""
from __future__ import annotations

import sys

import torch


def test_model(modelfile, modelclass, device=None, max_batches=None, config=None):
    import importlib
    from ..utils.recorder import Recorder
    cfg = dict(verbose=True, rank=0, size=1, mname=modelclass)
    if device:
        cfg["device"] = device
    cfg.update(config or {})
    model = getattr(importlib.import_module(modelfile), modelclass)(cfg)
    model.compile_iter_fns(sync_type="avg")
    fb = getattr(model, "file_batch_size", 128)
    recorder = Recorder(None, printFreq=max(1, 5120 // fb), modelname=modelclass, verbose=True, device=model.device)
    for epoch in range(model.n_epochs):
        model.epoch = epoch
        recorder.start_epoch()
        batch_i = 0
        n_train = model.data.n_batch_train if max_batches is None else min(max_batches, model.data.n_batch_train)
        while batch_i < n_train:
            for subb_i in range(model.n_subb):
                out = model.train_iter(batch_i, recorder)
            batch_i = out if isinstance(out, int) else batch_i + 1
            recorder.print_train_info(batch_i)
        recorder.clear_train_info()
        model.reset_iter("train")
        n_val = model.data.n_batch_val if max_batches is None else min(max_batches, model.data.n_batch_val)
        batch_j = 0
        stop = False
        while batch_j < n_val:
            for subb_i in range(model.n_subb):
                out = model.val_iter(batch_i, recorder)
                if out == "stop":
                    stop = True
                    break
                batch_j = out if isinstance(out, int) else batch_j + 1
            if stop:
                break
        model.reset_iter("val")
        recorder.print_val_info(batch_i)
        model.adjust_hyperp(epoch)
        if hasattr(model, "print_info"):
            model.print_info(recorder, True)
        recorder.end_epoch(batch_i, epoch)
        if stop or max_batches is not None:
            break
    model.cleanup()
    return recorder


if __name__ == "__main__":
    a = sys.argv[1:]
    test_model(a[0], a[1], a[2] if len(a) > 2 else None, int(a[3]) if len(a) > 3 else None)
