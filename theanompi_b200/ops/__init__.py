"""Functional op surface of the framework.

GPU: hand-written sm_100a kernels (``csrc/``) — tcgen05/TMEM/TMA GEMM with fused
bias/ReLU epilogues for FC and conv (implicit-GEMM via an NHWC im2col gather),
fused LRN / pool / dropout / softmax-xent / crop-mirror-normalise and the flat-arena
optimizer + collective kernels.  CPU: plain-torch reference (``reference.py``).
"""
from . import native, reference
from .functional import (add, advance_rng_step, batch_norm, compute_weight, conv2d_bias_act,
                         conv2d_group2_bias_act, crop_mirror_normalize, dropout, fork2,
                         linear_bias_act, lrn, pool2d, rng_state, seed_dropout,
                         softmax_xent)

__all__ = [
    "native", "reference", "conv2d_bias_act", "conv2d_group2_bias_act", "linear_bias_act",
    "pool2d", "lrn", "dropout", "softmax_xent", "crop_mirror_normalize", "compute_weight",
    "seed_dropout", "advance_rng_step", "rng_state", "batch_norm", "add", "fork2",
]
