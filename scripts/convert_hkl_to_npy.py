#!/usr/bin/env python
"""Convert the reference's ImageNet batch files (hickle ``.hkl``, uint8 ``c01b`` = channels × H × W × batch, written by the
reference's preprocessing, read in ``theanompi/models/data/proc_load_mpi.py:97``) to the ``.npy`` uint8 ``b01c`` (NHWC) files this
framework memory-maps.  Needs the ``hickle`` package on the machine that holds the data (it is not in this image; the loader
reads ``.hkl`` directly when it is).  Labels / ``img_mean.npy`` of the reference are used as they are.

    python scripts/convert_hkl_to_npy.py /data/imagenet/train_hkl_b256_b_128 /data/imagenet_npy/train_hkl_b256_b_128
"""
import glob
import os
import sys

import numpy as np


def main():
    src, dst = sys.argv[1], sys.argv[2]
    try:
        import hickle
    except ImportError:
        sys.exit("this converter needs the hickle package (pip install hickle) on the machine that holds the .hkl files")
    os.makedirs(dst, exist_ok=True)
    files = sorted(glob.glob(os.path.join(src, "*.hkl")))
    for i, f in enumerate(files):
        arr = np.asarray(hickle.load(f))
        if arr.ndim == 4 and arr.shape[0] in (1, 3):                 # c01b → b01c
            arr = np.transpose(arr, (3, 1, 2, 0))
        out = os.path.join(dst, os.path.splitext(os.path.basename(f))[0] + ".npy")
        np.save(out, np.ascontiguousarray(arr.astype(np.uint8, copy=False)))
        if i % 100 == 0:
            print("%d / %d  %s %s" % (i, len(files), out, arr.shape))


if __name__ == "__main__":
    main()
