"""Compute precision of the native (sm_100a) path.

``bf16`` (default)  bf16 activations and bf16 weight shadows feed ``tcgen05.mma kind::f16``; accumulation, master weights,
                    gradients and the optimizer are fp32.
``tf32``            fp32 storage end to end — activations, weights (no shadow copy), gradients — and GEMMs / convolutions on
                    ``tcgen05.mma kind::tf32``: the precision class of the reference, which computes in fp32 through
                    cuDNN / cuBLAS (``theanompi/models/layers2.py:380-388``, ``:927-929``).

Selected by ``TMPI_DTYPE`` / ``config['dtype']`` / :func:`set_precision` before the model is built.
"""
from __future__ import annotations

import os

import torch

_MODE = os.environ.get("TMPI_DTYPE", "bf16")
if _MODE not in ("bf16", "tf32"):
    raise ValueError("TMPI_DTYPE must be bf16 or tf32, got %r" % _MODE)


def set_precision(mode):
    global _MODE
    if mode not in ("bf16", "tf32"):
        raise ValueError("precision must be 'bf16' or 'tf32'")
    _MODE = mode


def precision():
    return _MODE


def tf32():
    return _MODE == "tf32"


def act_dtype():
    return torch.float32 if _MODE == "tf32" else torch.bfloat16
