"""Loader for the in-tree sm_100a extension ``_tmpi_native``.

The reference has no compiled code in the package: its GPU kernels are CUDA
strings JIT-compiled by libgpuarray/PyCUDA at run time
(``theanompi/lib/exchanger_strategy.py:163-174,296-307,637-646``).  Here every
kernel lives in ``theanompi_b200/csrc/*.cu`` and is compiled ahead of time
with ``nvcc -gencode arch=compute_100a,code=sm_100a`` into ONE shared object
that sits next to this file, so it travels to the GPU box with the source tree.

The binding layer is deliberately torch-free C++ (pybind11 + raw device
pointers + ``cudaStream_t`` handles): it builds in seconds, has no ABI coupling
to libtorch, and the Python wrappers in :mod:`theanompi_b200.ops` own all the
shape/dtype checking.

Policy: on a CUDA device a missing extension is a hard error (``require()``);
on CPU the pure-torch reference implementations are used.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
SO_NAME = "_tmpi_native.so"
SO_PATH = os.path.join(_PKG, SO_NAME)

_lock = threading.Lock()
_lib = None
_load_error = None


def _try_load():
    global _lib, _load_error
    if _lib is not None or _load_error is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(SO_PATH):
            _load_error = "extension not built: %s missing (run `python -m theanompi_b200.csrc.build`)" % SO_PATH
            return None
        try:
            spec = importlib.util.spec_from_file_location("_tmpi_native", SO_PATH)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            sys.modules.setdefault("_tmpi_native", mod)
            _lib = mod
        except Exception as e:  # pragma: no cover - depends on toolchain
            _load_error = "failed to load %s: %r" % (SO_PATH, e)
            return None
    return _lib


def available() -> bool:
    """True when the compiled extension can be imported (GPU or not)."""
    return _try_load() is not None


def lib():
    """Return the extension module or None."""
    return _try_load()


def load_error():
    _try_load()
    return _load_error


def require():
    """Return the extension module; raise loudly when it is missing.

    Called by every CUDA code path so a GPU box can never silently fall back
    to eager PyTorch for an op this framework claims as native.
    """
    m = _try_load()
    if m is None:
        raise RuntimeError(
            "theanompi_b200: native sm_100a extension is required on CUDA devices but "
            "is unavailable (%s)" % _load_error)
    return m


def launch_count() -> int:
    """Number of native kernel launches issued by this process so far."""
    m = _try_load()
    return int(m.launch_count()) if m is not None else 0


def reset_launch_count() -> None:
    m = _try_load()
    if m is not None:
        m.reset_launch_count()


def stream_ptr(t=None) -> int:
    """Raw ``cudaStream_t`` of torch's current stream for ``t``'s device."""
    import torch
    if t is not None and t.is_cuda:
        return torch.cuda.current_stream(t.device).cuda_stream
    return torch.cuda.current_stream().cuda_stream
