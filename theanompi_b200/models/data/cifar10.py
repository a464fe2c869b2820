"""CIFAR-10 dataset (ref ``theanompi/models/data/cifar10.py:23-213``): unpickle the five
python batches, 80/20 train/val split, list-of-batches with the same
``batch_data / extend_data / shuffle_data / shard_data`` API as ImageNet.  ``shuffle_data``
without a common seed is seeded by time·pid like the reference (the asynchronous rules
rely on different orders per worker).  Falls back to a synthetic set of the same shape
when the files are absent (no network in the build/bench environment).  Images are NHWC."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .utils import extend_data as _extend
from .utils import unpickle

data_path = os.environ.get("TMPI_CIFAR10_DIR", "./cifar-10-batches-py/")


class Cifar10_data(object):
    def __init__(self, verbose=False, synthetic=None, n_synthetic=2048, seed=0):
        self.verbose = verbose
        self.channels, self.width, self.height = 3, 32, 32
        self.n_class = 10
        self.batched = self.extended = False
        self.para_load = False
        self.loader = None
        if synthetic is None:
            synthetic = not os.path.isdir(data_path)
        self.synthetic = synthetic
        self.get_data(n_synthetic, seed)

    def get_data(self, n_synthetic=2048, seed=0):
        if self.synthetic:
            rs = np.random.RandomState(seed)
            # class-dependent means make the synthetic task learnable (used by convergence tests)
            labels = rs.randint(0, self.n_class, n_synthetic).astype(np.int64)
            centers = rs.randn(self.n_class, 1, 1, 3).astype(np.float32) * 40 + 128
            img = centers[labels] + rs.randn(n_synthetic, 32, 32, 3).astype(np.float32) * 20
            img = np.clip(img, 0, 255).astype(np.float32)
        else:
            xs, ys = [], []
            for i in range(1, 6):
                d = unpickle(os.path.join(data_path, "data_batch_%d" % i))
                xs.append(np.asarray(d["data"]).reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1))
                ys.append(np.asarray(d["labels"]))
            img = np.concatenate(xs).astype(np.float32)
            labels = np.concatenate(ys).astype(np.int64)
        n = len(img)
        n_train = int(n * 0.8)
        self.img_mean = img[:n_train].mean(axis=0)
        self.rawdata = [img[:n_train], labels[:n_train], img[n_train:], labels[n_train:], self.img_mean, None]

    def batch_data(self, file_batch_size):
        if self.batched:
            return
        def split(x, y):
            nb = len(x) // file_batch_size
            return ([x[i * file_batch_size:(i + 1) * file_batch_size] for i in range(nb)],
                    [y[i * file_batch_size:(i + 1) * file_batch_size] for i in range(nb)])
        self.train_img, self.train_labels = split(self.rawdata[0], self.rawdata[1])
        self.val_img, self.val_labels = split(self.rawdata[2], self.rawdata[3])
        self.n_batch_train, self.n_batch_val = len(self.train_img), len(self.val_img)
        if self.verbose:
            print("train on %d batches, val on %d batches" % (self.n_batch_train, self.n_batch_val))
        self.batched = True

    def extend_data(self, rank, size):
        if self.extended:
            return
        self.train_img_ext, self.train_labels_ext = _extend(rank, size, self.train_img, self.train_labels)
        self.val_img_ext, self.val_labels_ext = _extend(rank, size, self.val_img, self.val_labels)
        self.n_batch_train, self.n_batch_val = len(self.train_img_ext), len(self.val_img_ext)
        self.extended = True

    def shuffle_data(self, mode, common_seed=None):
        if mode == "train":
            seed = common_seed if common_seed is not None else (int(time.time() * 1000) * os.getpid()) % (2 ** 31)
            rs = np.random.RandomState(seed)
            idx = rs.permutation(len(self.train_img_ext))
            self.train_img_shuffle = [self.train_img_ext[i] for i in idx]
            self.train_labels_shuffle = [self.train_labels_ext[i] for i in idx]
        else:
            self.val_img_shuffle, self.val_labels_shuffle = self.val_img_ext, self.val_labels_ext

    def shard_data(self, mode, rank, size):
        if mode == "train":
            self.train_img_shard = self.train_img_shuffle[rank::size]
            self.train_labels_shard = self.train_labels_shuffle[rank::size]
            self.n_batch_train = len(self.train_img_shard)
        else:
            self.val_img_shard = self.val_img_shuffle[rank::size]
            self.val_labels_shard = self.val_labels_shuffle[rank::size]
            self.n_batch_val = len(self.val_img_shard)

    def load_batch(self, item, mode, model):
        """In-memory batch → pinned → device (no loader process needed for 3 MB batches)."""
        if not model.cuda:
            return torch.from_numpy(np.ascontiguousarray(item))
        # persistent page-locked staging ring (allocating pinned memory per batch costs milliseconds); a slot is reused only
        # after the H2D copy that read it has completed
        shape = tuple(item.shape)
        ring = getattr(self, "_ring", None)
        if ring is None or ring["shape"] != shape:
            ring = self._ring = dict(shape=shape, i=0, ev=[None, None],
                                     buf=[torch.empty(shape, dtype=torch.float32, pin_memory=True) for _ in range(2)])
        k = ring["i"]; ring["i"] = 1 - k
        if ring["ev"][k] is not None:
            ring["ev"][k].synchronize()
        ring["buf"][k].numpy()[...] = item
        t = ring["buf"][k].to(model.device, non_blocking=True)
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream(model.device)); ring["ev"][k] = ev
        return t
