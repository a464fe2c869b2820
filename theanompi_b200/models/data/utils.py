"""Data helpers (ref ``theanompi/models/data/utils.py``): ``unpickle`` (``:3-12``),
``get_bad_list``/``extend_data`` — pad the file list to a multiple of the world
size (``:15-40``), ``get_rand3d`` and the CPU ``crop_and_mirror`` (``:42-129``).

Images are NHWC ``[N, H, W, C]`` here (the reference used c01b).  The CPU
``crop_and_mirror`` is kept as the host fallback and as ground truth for the device
kernel (``csrc/data_kernels.cu``), which does normalise + crop + mirror + bf16 cast in
one pass after the pinned H2D copy.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

_COPY_POOL = None


def parallel_copyto(out, src, threads=None, min_bytes=4 << 20):
    """``np.copyto(out, src)`` split over a few threads along the first axis.  One core moves a 25 MB file batch out of the
    page cache at 4–6 GB/s (≈ 4–6 ms) — slower than a B200 consumes it (AlexNet-128b: one batch per 1.6 ms), so the host copy of
    the loader can be spread over ``TMPI_LOADER_THREADS`` cores (default 1; numpy releases the GIL inside each slice copy).  On
    the 8-vCPU build sandbox this does not help (one core already saturates its memory bandwidth, ``profiles/loader_host.md``);
    it is meant for real hosts."""
    global _COPY_POOL
    n = int(threads or os.environ.get("TMPI_LOADER_THREADS", "1"))
    n = max(1, min(n, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else n, int(out.shape[0]) if out.ndim else 1))
    if n == 1 or out.nbytes < min_bytes:
        np.copyto(out, src)
        return
    if _COPY_POOL is None or _COPY_POOL._max_workers < n:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=n, thread_name_prefix="tmpi-copy")
    rows = int(out.shape[0])
    step = (rows + n - 1) // n
    futs = [_COPY_POOL.submit(np.copyto, out[a:a + step], src[a:a + step]) for a in range(0, rows, step)]
    for f in futs:
        f.result()


def unpickle(path):
    with open(path, "rb") as f:
        try:
            return pickle.load(f, encoding="latin1")
        except TypeError:
            return pickle.load(f)


def get_bad_list(n_batches, commsize):
    bad_left = n_batches % commsize
    return [n_batches - (bad + 1) for bad in range(bad_left)]


def extend_data(rank, size, img_batches, label_batches, verbose=False):
    """Repeat trailing batches so ``len % size == 0`` (every rank gets the same
    number of file batches)."""
    _img = list(img_batches)
    n_files = len(_img)
    _lab = list(label_batches[:n_files])
    bad_left_list = get_bad_list(n_files, size)
    need = (size - len(bad_left_list)) % size
    if need != 0:
        _img.extend(_img[-need:])
        _lab.extend(_lab[-need:])
    assert len(_img) % size == 0
    if rank == 0 and verbose:
        print("rank%d: bad list is %s, extended to %d" % (rank, str(bad_left_list), len(_img)))
    return _img, _lab


def get_rand3d(rand_crop, mode, rs=None):
    """Three uniforms in [0,1): (y-offset, x-offset, flip); 0.5/0.5/0 for val."""
    if not rand_crop or mode == "val":
        return np.float32([0.5, 0.5, 0])
    rs = rs or np.random
    return np.float32(rs.rand(3))


def draw_crops(n, in_hw, out_hw, mode, rand_crop=True, batch_crop_mirror=False, rs=None):
    """Per-image (y0, x0) offsets and flip flags; centre crop / no flip for val."""
    H, W = in_hw
    ch, cw = out_hw
    rs = rs or np.random
    if mode == "val" or not rand_crop:
        offs = np.tile(np.int32([[(H - ch) // 2, (W - cw) // 2]]), (n, 1))
        return offs, np.zeros(n, dtype=np.uint8)
    if batch_crop_mirror:
        r = get_rand3d(True, mode, rs)
        offs = np.tile(np.int32([[int(r[0] * (H - ch + 1)), int(r[1] * (W - cw + 1))]]), (n, 1))
        flips = np.full(n, int(r[2] > 0.5), dtype=np.uint8)
        return offs, flips
    oy = rs.randint(0, H - ch + 1, n)
    ox = rs.randint(0, W - cw + 1, n)
    flips = (rs.rand(n) > 0.5).astype(np.uint8)
    return np.stack([oy, ox], 1).astype(np.int32), flips


def crop_and_mirror(data, mode, rand_crop, flag_batch, cropsize, rs=None):
    """Host reference: ``data`` is NHWC float/uint8; returns NHWC ``cropsize²`` crops."""
    n, H, W, C = data.shape
    offs, flips = draw_crops(n, (H, W), (cropsize, cropsize), mode, rand_crop, flag_batch, rs)
    out = np.empty((n, cropsize, cropsize, C), dtype=data.dtype)
    for i in range(n):
        y0, x0 = offs[i]
        patch = data[i, y0:y0 + cropsize, x0:x0 + cropsize, :]
        out[i] = patch[:, ::-1, :] if flips[i] else patch
    return np.ascontiguousarray(out)
