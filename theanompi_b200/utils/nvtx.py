"""NVTX ranges for Nsight timelines (the reference has no profiler hooks at all).
Enabled with ``TMPI_NVTX=1``; zero cost otherwise."""
from __future__ import annotations

import contextlib
import os

ENABLED = os.environ.get("TMPI_NVTX", "0") == "1"


@contextlib.contextmanager
def range(name):
    if not ENABLED:
        yield
        return
    import torch
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()
