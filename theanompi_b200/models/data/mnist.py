"""MNIST (ref ``theanompi/models/data/mnist.py``): infinite shuffled minibatch generators
for the GAN models.  The reference downloads ``mnist.pkl.gz``; without network access a
synthetic 28×28 digit-like set (class-dependent blobs) of the same shape is generated."""
from __future__ import annotations

import gzip
import os
import pickle

import numpy as np

path = os.environ.get("TMPI_MNIST_PATH", "./mnist.pkl.gz")


class MNIST_data(object):
    def __init__(self, verbose=False, n_synthetic=4096, seed=0):
        self.verbose = verbose
        self.channels, self.width, self.height, self.n_class = 1, 28, 28, 10
        if os.path.exists(path):
            with gzip.open(path, "rb") as f:
                tr, va, te = pickle.load(f, encoding="latin1")
            self.train_x, self.train_y = tr[0].reshape(-1, 28, 28, 1).astype(np.float32), tr[1].astype(np.int64)
            self.val_x, self.val_y = va[0].reshape(-1, 28, 28, 1).astype(np.float32), va[1].astype(np.int64)
        else:
            rs = np.random.RandomState(seed)
            y = rs.randint(0, 10, n_synthetic).astype(np.int64)
            yy, xx = np.mgrid[0:28, 0:28].astype(np.float32)
            cx = 6 + 1.8 * (y % 5)[:, None, None]; cy = 8 + 10 * (y // 5)[:, None, None]
            x = np.exp(-((xx[None] - cx) ** 2 + (yy[None] - cy) ** 2) / 18.0) + 0.05 * rs.rand(n_synthetic, 28, 28)
            x = np.clip(x, 0, 1).astype(np.float32)[..., None]
            n = int(0.9 * n_synthetic)
            self.train_x, self.train_y, self.val_x, self.val_y = x[:n], y[:n], x[n:], y[n:]
        self.n_batch_train = self.n_batch_val = None
        self.para_load = False

    def batch_data(self, batch_size):
        self.batch_size = batch_size
        self.n_batch_train = len(self.train_x) // batch_size
        self.n_batch_val = max(1, len(self.val_x) // batch_size)

    def extend_data(self, rank, size):
        pass

    def shuffle_data(self, mode="train", common_seed=None):
        pass

    def shard_data(self, mode, rank, size):
        self.rank, self.size = rank, size
        if mode == "train":
            self.n_batch_train = max(1, len(self.train_x) // self.batch_size // size)

    def iterate(self, mode="train", shuffle=True, seed=None, forever=True):
        """Infinite generator of (x[B,28,28,1], y[B]) minibatches (ref ``mnist.py:120-156``)."""
        x, y = (self.train_x, self.train_y) if mode == "train" else (self.val_x, self.val_y)
        rs = np.random.RandomState(seed)
        B = self.batch_size
        while True:
            idx = rs.permutation(len(x)) if shuffle else np.arange(len(x))
            if len(idx) < B:                        # tiny (synthetic / test) sets: tile up to one batch
                idx = np.resize(idx, B)
            for s in range(0, len(idx) - B + 1, B):
                sel = idx[s:s + B]
                yield x[sel], y[sel]
            if not forever:
                break
