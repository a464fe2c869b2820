"""Wasserstein GAN / LSGAN on MNIST (ref ``lasagne_model_zoo/wgan.py``, ``lsgan.py``).

Contract quirks kept from the reference: the exchanged ``params`` are the **critic**
parameters only (``wgan.py:142``); ``train_iter`` runs 50 (first 5 and every 100th
generator update) or 2 critic steps with weight clipping ±0.01, then one generator step,
and *returns the advanced count* (``:240-270``); ``val_iter`` records (critic score,
generator score, 0); ``print_info`` plots samples + score curves (``:287-312``); own
``save/load`` (npz); lr decays linearly to zero over the second half (``:314-320``).
RMSProp as in the reference (``wgan.py:18-59``).  DCGAN-style generator / critic
(``:61-110``) built from torch modules.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ..torch_base import TorchModelBase, tag_module_params

num_epochs = 100
epochsize = 100
batchsize = 64
initial_eta = 5e-5
clip = 0.01


def build_generator(nz=100, out_ch=1, size=28):
    s4 = size // 4
    return nn.Sequential(nn.Linear(nz, 1024), nn.BatchNorm1d(1024), nn.ReLU(True),
                         nn.Linear(1024, 128 * s4 * s4), nn.BatchNorm1d(128 * s4 * s4), nn.ReLU(True),
                         nn.Unflatten(1, (128, s4, s4)),
                         nn.ConvTranspose2d(128, 64, 5, 2, 2, output_padding=1), nn.BatchNorm2d(64), nn.ReLU(True),
                         nn.ConvTranspose2d(64, out_ch, 5, 2, 2, output_padding=1), nn.Sigmoid())


def build_critic(in_ch=1, size=28):
    s4 = size // 4
    return nn.Sequential(nn.Conv2d(in_ch, 64, 5, 2, 2), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(64, 128, 5, 2, 2), nn.BatchNorm2d(128), nn.LeakyReLU(0.2, True),
                         nn.Flatten(), nn.Linear(128 * s4 * s4, 1024), nn.BatchNorm1d(1024), nn.LeakyReLU(0.2, True),
                         nn.Linear(1024, 1))


class WGAN(TorchModelBase):
    loss_kind = "wgan"
    n_epochs = num_epochs
    batch_size = file_batch_size = batchsize
    learning_rate = initial_eta
    image_size, image_ch = 28, 1

    def __init__(self, config):
        super().__init__(config)
        self.name = "Wasserstein_GAN" if self.loss_kind == "wgan" else "LSGAN"
        torch.manual_seed(1234)
        self.n_epochs = config.get("n_epochs", self.n_epochs)
        self.epochsize = config.get("epochsize", epochsize)
        self.data = self.make_data(config)
        self.n_subb = 1
        self.generator = build_generator(100, self.image_ch, self.image_size).to(self.device)
        self.critic = build_critic(self.image_ch, self.image_size).to(self.device)
        self.config["_arena_shadow"] = False
        cparams, ctypes = tag_module_params(self.critic)
        for p in cparams:                       # GAN critics: everything (BN included) follows the same rule
            p.pname = "W" if p.dim() > 1 else "b"
        self.finalize(cparams, ctypes, (self.batch_size, self.image_size, self.image_size, self.image_ch))
        for p in self.params:
            p.grad, p.shadow = p.gbuf, None
        self.critic_params = self.params
        self.generator_params = [p for p in self.generator.parameters()]
        self.generator_updates = 0
        self.critic_scores, self.generator_scores, self.c_list, self.g_list = [], [], [], []
        self.current_info = None
        self.init_view = False
        self._train_gen = self.data.iterate("train", seed=1234 + self.rank)
        self._val_gen = self.data.iterate("val", shuffle=False)
        self.data.n_batch_train = self.epochsize
        self.data.n_batch_val = 1

    def make_data(self, config):
        from ..data.mnist import MNIST_data
        d = MNIST_data(self.verbose, **config.get("data_kwargs", {}))
        d.batch_data(self.batch_size)
        return d

    # ---- losses
    def _critic_loss(self, real, fake):
        if self.loss_kind == "wgan":
            return self.critic(fake).mean() - self.critic(real).mean()
        return 0.5 * ((self.critic(real) - 1) ** 2).mean() + 0.5 * (self.critic(fake) ** 2).mean()

    def _gen_loss(self, fake):
        if self.loss_kind == "wgan":
            return -self.critic(fake).mean()
        return 0.5 * ((self.critic(fake) - 1) ** 2).mean()

    def compile_iter_fns(self, sync_type="avg", **kw):
        self.sync_type = "avg"
        self.opt_c = torch.optim.RMSprop(self.critic_params, lr=self.learning_rate)
        self.opt_g = torch.optim.RMSprop(self.generator_params, lr=self.learning_rate)
        self.vels, self.vels2 = [], []
        self.train_iter_fn = self.val_iter_fn = None

    def _batch(self, gen):
        x, _ = next(gen)
        return torch.from_numpy(np.ascontiguousarray(x)).to(self.device).permute(0, 3, 1, 2).float()

    def _noise(self, n):
        return torch.rand(n, 100, device=self.device)

    def critic_train_fn(self, real):
        for p in self.critic_params:
            p.grad = p.gbuf
        self.arena.G.zero_()
        with torch.no_grad():
            fake = self.generator(self._noise(real.shape[0]))
        loss = self._critic_loss(real, fake)
        loss.backward()
        self.opt_c.step()
        return -loss.detach() if self.loss_kind == "wgan" else loss.detach()

    def critic_clip_fn(self):
        if self.loss_kind == "wgan":
            with torch.no_grad():
                self.arena.W.clamp_(-clip, clip)

    def generator_train_fn(self):
        self.opt_g.zero_grad(set_to_none=True)
        loss = self._gen_loss(self.generator(self._noise(self.batch_size)))
        loss.backward()
        self.opt_g.step()
        self.arena.G.zero_()
        return loss.detach()

    def train_iter(self, count, recorder):
        if self.loss_kind == "wgan":
            critic_runs = 50 if (self.generator_updates < 5 or self.generator_updates % 100 == 0) else 2
            critic_runs = self.config.get("critic_runs", critic_runs)
        else:
            critic_runs = 1
        scores = []
        recorder.start()
        self.critic.train(); self.generator.train()
        for _ in range(critic_runs):
            scores.append(self.critic_train_fn(self._batch(self._train_gen)))
            self.critic_clip_fn()
            count += 1
        g_score = self.generator_train_fn()
        self.critic_scores.extend(float(s) for s in scores)
        self.generator_scores.append(float(g_score))
        self.generator_updates += 1
        recorder.train_error(count, sum(scores) / len(scores), g_score)
        recorder.end("calc")
        return count

    def val_iter(self, count, recorder):
        self.critic.eval(); self.generator.eval()
        with torch.no_grad():
            real = self._batch(self._val_gen)
            fake = self.generator(self._noise(real.shape[0]))
            c, g = self._critic_loss(real, fake), self._gen_loss(fake)
        recorder.val_error(count, -c if self.loss_kind == "wgan" else c, g, 0)

    def reset_iter(self, *args, **kwargs):
        pass

    def print_info(self, recorder, verbose=True):
        if not self.generator_scores:
            return
        g_, c_ = float(np.mean(self.generator_scores)), float(np.mean(self.critic_scores))
        self.g_list.append(g_); self.c_list.append(c_)
        if verbose:
            print("\nEpoch %d\n  generator score:\t\t%s\n  %s:\t\t%s" % (self.epoch, g_, "Wasserstein distance" if self.loss_kind == "wgan" else "critic loss", c_))
        self.critic_scores[:] = []; self.generator_scores[:] = []
        if verbose and self.config.get("plot", False):
            with torch.no_grad():
                self.generator.eval()
                s = self.generator(self._noise(42)).float().cpu().numpy()
            img = s[:, 0].reshape(6, 7, self.image_size, self.image_size).transpose(0, 2, 1, 3).reshape(6 * self.image_size, 7 * self.image_size)
            if not self.init_view:
                self.init_view = True
                recorder.plot_init(name="scores", save=True); recorder.plot_init(name="sample", save=True)
            recorder.plot(name="sample", image=img, cmap="gray")
            recorder.plot(name="scores", lines=[(list(range(len(self.c_list))), self.c_list, "critic"),
                                                (list(range(len(self.g_list))), self.g_list, "generator")])

    def adjust_hyperp(self, epoch):
        if epoch >= self.n_epochs // 2:
            progress = float(epoch) / self.n_epochs
            lr = self.learning_rate * 2 * (1 - progress)
            self.shared_lr.set_value(lr)
            for opt in (self.opt_c, self.opt_g):
                for g in opt.param_groups:
                    g["lr"] = lr

    def cleanup(self):
        pass

    def save(self, path):
        import os
        os.makedirs(path, exist_ok=True)
        np.savez(os.path.join(path, "%s_gen_%d.npz" % (self.name, self.epoch)), *[p.detach().cpu().numpy() for p in self.generator_params])
        np.savez(os.path.join(path, "%s_crit_%d.npz" % (self.name, self.epoch)), *[p.detach().cpu().numpy() for p in self.critic_params])

    def load(self, path, epoch):
        import os
        for params, tag in ((self.generator_params, "gen"), (self.critic_params, "crit")):
            with np.load(os.path.join(path, "%s_%s_%d.npz" % (self.name, tag, epoch))) as f:
                with torch.no_grad():
                    for p, k in zip(params, sorted(f.files, key=lambda s: int(s.split("_")[1]))):
                        p.copy_(torch.from_numpy(f[k]).to(p.device))
