"""Weight I/O, checkpoint/resume and the model-contract checker.

Reference: ``theanompi/lib/helper_funcs.py`` — per-layer ``.npy`` save/load by
attribute name (``:33-143``), momentum save/load (``:147-161``), ``check_model``
(``:163-205``), ``check_model_cdd`` (``:207-227``), ``save_model`` (``:230-260``).

Reference gaps closed here: ``save_model`` wrote the
learning rate as a constant 0, never saved momentum, and no code path ever loaded a
snapshot.  ``save_checkpoint`` / ``load_checkpoint`` store weights + momentum + lr +
epoch + recorder curves in one file, and every worker accepts ``resume=<path>``.
"""
from __future__ import annotations

import glob
import os
import pickle

import numpy as np
import torch

_ATTRS = ("W", "W0", "W1", "b", "b0", "b1", "gamma", "beta")


def bufint(t):
    """Device pointer of a tensor (the reference wrapped a gpuarray in a
    ``memoryview`` for CUDA-aware MPI, ``helper_funcs.py:19-23``).  Here the raw
    pointer is what the peer-memory registry and the kernels consume."""
    return int(t.data_ptr())


def dtype_to_mpi(t):
    """numpy/torch dtype → wire-type name (ref ``helper_funcs.py:25-29``)."""
    name = str(t).replace("torch.", "")
    return {"float32": "FLOAT", "float16": "HALF", "bfloat16": "BFLOAT16",
            "float64": "DOUBLE", "int32": "INT", "int64": "LONG"}.get(name, name.upper())


# --------------------------------------------------------------------------- per-layer npy
def save_weights(layers, weights_dir, epoch):
    os.makedirs(weights_dir, exist_ok=True)
    for idx, layer in enumerate(layers):
        for a in _ATTRS:
            if hasattr(layer, a) and hasattr(getattr(layer, a), "save_weight"):
                getattr(layer, a).save_weight(weights_dir, "%s_%d_%s" % (a, idx, epoch))


def load_weights(layers, weights_dir, epoch, l_range=None):
    for idx, layer in enumerate(layers):
        if l_range is not None and idx not in l_range:
            continue
        for a in _ATTRS:
            if hasattr(layer, a) and hasattr(getattr(layer, a), "load_weight"):
                getattr(layer, a).load_weight(weights_dir, "%s_%d_%s" % (a, idx, epoch))


def collect_weight_path(layers, weights_dir, epoch):
    paths = []
    for idx, layer in enumerate(layers):
        for a in _ATTRS:
            if hasattr(layer, a):
                paths.append(os.path.join(weights_dir, "%s_%d_%s.npy" % (a, idx, epoch)))
    return paths


def load_weights_from_memory(layers, arrays):
    """``arrays``: list in the order produced by :func:`collect_weight_path`."""
    it = iter(arrays)
    with torch.no_grad():
        for layer in layers:
            for a in _ATTRS:
                if hasattr(layer, a):
                    w = getattr(layer, a)
                    arr = next(it)
                    w.val.copy_(torch.as_tensor(np.asarray(arr)).to(w.val.device))
                    sh = getattr(w.val, "shadow", None)
                    if sh is not None:
                        sh.copy_(w.val)


def save_momentums(vels, weights_dir, epoch):
    os.makedirs(weights_dir, exist_ok=True)
    for ind, v in enumerate(vels):
        np.save(os.path.join(weights_dir, "mom_%d_%s.npy" % (ind, epoch)), v.detach().float().cpu().numpy())


def load_momentums(vels, weights_dir, epoch):
    with torch.no_grad():
        for ind, v in enumerate(vels):
            arr = np.load(os.path.join(weights_dir, "mom_%d_%s.npy" % (ind, epoch)))
            v.copy_(torch.from_numpy(arr).to(v.device))


# --------------------------------------------------------------------------- contract
_CONTRACT_MSG = (
    "Model def lacks some attributes and/or methods\n"
    "attributes include: params (list of torch tensors), data, epoch (initialized to 0),\n"
    "                    n_epochs (max epochs), n_subb (sub batches per minibatch, default 1)\n"
    "methods include: compile_iter_fns, train_iter, val_iter, reset_iter, adjust_hyperp, cleanup\n")


def check_model(model):
    """Duck-typed model contract (ref ``helper_funcs.py:163-205``)."""
    try:
        assert hasattr(model, "params") and isinstance(model.params, list) and len(model.params) > 0
        assert isinstance(model.params[0], torch.Tensor)
        for a in ("data", "epoch", "n_epochs", "n_subb"):
            assert hasattr(model, a), a
        for m in ("compile_iter_fns", "train_iter", "val_iter", "reset_iter", "adjust_hyperp", "cleanup"):
            assert callable(getattr(model, m, None)), m
    except AssertionError:
        print(_CONTRACT_MSG)
        raise


def check_model_cdd(model):
    """Ensure ``vels``/``vels2`` lists exist (ref ``helper_funcs.py:207-227``)."""
    if isinstance(getattr(model, "vels", None), list) and isinstance(getattr(model, "vels2", None), list):
        return
    arena = getattr(model, "arena", None)
    if arena is not None:
        model.vels, model.vels2 = arena.views("G"), arena.views("R")
    else:
        model.vels = [torch.zeros_like(p) for p in model.params]
        model.vels2 = [torch.zeros_like(p) for p in model.params]


# --------------------------------------------------------------------------- snapshots
def save_model(model, path, verbose, recorder=None):
    """Reference-compatible snapshot (ref ``:230-260``) + a full resumable checkpoint."""
    os.makedirs(path, exist_ok=True)
    if hasattr(model, "save") and callable(model.save):
        model.save(path)
    else:
        layers = getattr(model, "layers", None)
        if layers:
            save_weights(layers, path, model.epoch)
        else:
            with open(os.path.join(path, "%sparams_%d.pkl" % (getattr(model, "name", "model"), model.epoch)), "wb") as f:
                pickle.dump([p.detach().cpu() for p in model.params], f, protocol=pickle.HIGHEST_PROTOCOL)
        lr = model.shared_lr.get_value() if hasattr(model, "shared_lr") else 0.0
        np.save(os.path.join(path, "lr_%d.npy" % model.epoch), np.float32(lr))
    save_checkpoint(model, os.path.join(path, "ckpt_%d.pt" % model.epoch), recorder=recorder)
    if verbose:
        print("\nweights saved at epoch %d" % model.epoch)
    try:
        with open(os.path.join(path, "val_info.txt"), "a") as f:
            f.write("\nepoch: {} val_info {}:".format(model.epoch, getattr(model, "current_info", None)))
    except OSError:
        pass


def save_checkpoint(model, filename, recorder=None, extra=None):
    sd = {"epoch": int(model.epoch), "name": getattr(model, "name", "model")}
    if hasattr(model, "shared_lr"):
        sd["lr"] = float(model.shared_lr.get_value())
    arena = getattr(model, "arena", None)
    if arena is not None:
        sd["arena"] = arena.state_dict()
    else:
        sd["params"] = [p.detach().cpu() for p in model.params]
    if hasattr(model, "extra_state"):
        sd["extra_state"] = model.extra_state()
    if recorder is not None:
        sd["recorder"] = recorder.info_dict
    if extra:
        sd["extra"] = extra
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    tmp = filename + ".tmp"
    torch.save(sd, tmp)
    os.replace(tmp, filename)
    return filename


def load_checkpoint(model, filename, recorder=None):
    """Restore weights, momentum, lr, epoch (and curves).  Returns the epoch to
    resume FROM (saved epoch + 1)."""
    sd = torch.load(filename, map_location="cpu", weights_only=False)
    arena = getattr(model, "arena", None)
    if arena is not None and "arena" in sd:
        arena.load_state_dict(sd["arena"])
    elif "params" in sd:
        with torch.no_grad():
            for p, q in zip(model.params, sd["params"]):
                p.copy_(q.to(p.device))
    if "lr" in sd and hasattr(model, "shared_lr"):
        model.shared_lr.set_value(sd["lr"])
    if "extra_state" in sd and hasattr(model, "load_extra_state"):
        model.load_extra_state(sd["extra_state"])
    model.epoch = int(sd["epoch"])
    if recorder is not None and "recorder" in sd:
        for k, v in sd["recorder"].items():
            recorder.info_dict[k] = list(v)
    return model.epoch + 1


def latest_checkpoint(path):
    files = glob.glob(os.path.join(path, "ckpt_*.pt"))
    if not files:
        return None
    return max(files, key=lambda f: int(os.path.basename(f)[5:-3]))
