"""LSTM sentiment classifier on IMDB (ref ``theanompi/models/lstm.py`` — the Theano tutorial
LSTM: 128 units, maxlen 500, batch 16, mean-pooling over time, Adadelta ``:284-342`` — and
its model-contract adapter ``lstm_theanompi_outdated.py`` with rank-sharded ``IMDB_Data``
``:75-94`` and early stopping through ``val_iter`` returning ``'stop'``).

The recurrent cell is ``torch.nn.LSTM`` (cuDNN); sequences are never split across
devices (the reference has no sequence parallelism).  Without the IMDB pickle
a synthetic corpus with class-dependent token statistics is generated.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from .torch_base import TorchModelBase

dim_proj = 128
maxlen = 500
n_words = 10000
batch_size = 16
patience = 10
max_epochs = 100


class IMDB_Data(object):
    def __init__(self, rank=0, size=1, n_synthetic=512, seed=0, maxlen=maxlen, n_words=n_words):
        path = os.environ.get("TMPI_IMDB_PATH", "./imdb.pkl")
        rs = np.random.RandomState(seed)
        if os.path.exists(path):
            import pickle
            with open(path, "rb") as f:
                tr = pickle.load(f); te = pickle.load(f)
            xs = [np.asarray([min(w, n_words - 1) for w in s][:maxlen]) for s in tr[0]]
            ys = np.asarray(tr[1])
        else:
            ys = rs.randint(0, 2, n_synthetic)
            xs = []
            for y in ys:
                L = rs.randint(20, 80)
                base = rs.randint(2, n_words, L)
                marks = rs.rand(L) < 0.3
                base[marks] = (2 + y * 50 + rs.randint(0, 50, marks.sum()))        # sentiment-bearing tokens
                xs.append(base)
        n = len(xs)
        nv = max(1, n // 10)
        self.train = (xs[nv:][rank::size], ys[nv:][rank::size])        # shard by rank (ref :75-94)
        self.valid = (xs[:nv], ys[:nv])
        self.n_words, self.maxlen = n_words, maxlen

    def batches(self, split, bs, shuffle, seed=None):
        xs, ys = self.train if split == "train" else self.valid
        idx = np.random.RandomState(seed).permutation(len(xs)) if shuffle else np.arange(len(xs))
        for s in range(0, len(xs) - bs + 1, bs):
            sel = idx[s:s + bs]
            L = max(len(xs[i]) for i in sel)
            x = np.zeros((bs, L), dtype=np.int64); m = np.zeros((bs, L), dtype=np.float32)
            for j, i in enumerate(sel):
                x[j, :len(xs[i])] = xs[i]; m[j, :len(xs[i])] = 1
            yield x, m, np.asarray([ys[i] for i in sel], dtype=np.int64)


class LSTMNet(nn.Module):
    def __init__(self, n_words, dim, n_out=2):
        super().__init__()
        self.emb = nn.Embedding(n_words, dim)
        self.lstm = nn.LSTM(dim, dim, batch_first=True)
        self.drop = nn.Dropout(0.5)
        self.out = nn.Linear(dim, n_out)

    def forward(self, x, mask):
        h, _ = self.lstm(self.emb(x))
        pooled = (h * mask[..., None]).sum(1) / mask.sum(1, keepdim=True).clamp_min(1)    # mean pooling (ref :217-253)
        return self.out(self.drop(pooled))


class LSTM(TorchModelBase):
    n_epochs = max_epochs
    batch_size = file_batch_size = batch_size
    learning_rate = 1.0
    autocast = False

    def __init__(self, config):
        super().__init__(config)
        self.name = "LSTM"
        torch.manual_seed(123)
        self.n_epochs = config.get("n_epochs", self.n_epochs)
        self.data = IMDB_Data(self.rank, self.size, **config.get("data_kwargs", {}))
        self.n_subb = 1
        self.config["_arena_shadow"] = False
        self.finalize_torch(LSTMNet(self.data.n_words, config.get("dim_proj", dim_proj)), (self.batch_size, 1))
        self.data.n_batch_train = len(self.data.train[0]) // self.batch_size
        self.data.n_batch_val = max(1, len(self.data.valid[0]) // self.batch_size)
        self.best_err, self.bad_counter, self.patience = 1.0, 0, config.get("patience", patience)
        self._val_errs = []
        self._train_it = None

    def make_torch_optimizer(self, params):
        return torch.optim.Adadelta(params, lr=1.0, rho=0.95, eps=1e-6)

    def compile_iter_fns(self, sync_type="avg", **kw):
        self.sync_type = "avg"
        self.torch_opt = self.make_torch_optimizer(self.params)
        self.vels, self.vels2 = [], []

    def _to(self, x, m, y):
        d = self.device
        return torch.from_numpy(x).to(d), torch.from_numpy(m).to(d), torch.from_numpy(y).to(d)

    def train_iter(self, count, recorder):
        if self._train_it is None:
            self._train_it = self.data.batches("train", self.batch_size, True, seed=self.epoch)
        try:
            x, m, y = next(self._train_it)
        except StopIteration:
            self._train_it = self.data.batches("train", self.batch_size, True, seed=self.epoch + 1000)
            x, m, y = next(self._train_it)
        recorder.start()
        x, m, y = self._to(x, m, y)
        self.module.train()
        self.arena.G.zero_()
        for p in self.params:
            p.grad = p.gbuf
        logits = self.module(x, m)
        cost = nn.functional.cross_entropy(logits, y)
        cost.backward()
        self.torch_opt.step()
        err = (logits.argmax(1) != y).float().mean()
        recorder.train_error(count, cost.detach(), err)
        recorder.end("calc")

    def val_iter(self, count, recorder):
        """One full validation pass; returns ``'stop'`` when patience runs out (the early-stop
        protocol ``BSP_run`` understands, ``worker.py:118-126``)."""
        self.module.eval()
        errs, costs = [], []
        with torch.no_grad():
            for x, m, y in self.data.batches("valid", min(self.batch_size, len(self.data.valid[0])), False):
                x, m, y = self._to(x, m, y)
                lg = self.module(x, m)
                costs.append(float(nn.functional.cross_entropy(lg, y)))
                errs.append(float((lg.argmax(1) != y).float().mean()))
        e, c = float(np.mean(errs)), float(np.mean(costs))
        recorder.val_error(count, c, e, 0)
        if e < self.best_err:
            self.best_err, self.bad_counter = e, 0
        else:
            self.bad_counter += 1
            if self.bad_counter > self.patience:
                return "stop"
        return self.data.n_batch_val            # one call covers the whole validation set

    def reset_iter(self, mode):
        if mode == "train":
            self._train_it = None

    def adjust_hyperp(self, epoch):
        pass

    def cleanup(self):
        pass
