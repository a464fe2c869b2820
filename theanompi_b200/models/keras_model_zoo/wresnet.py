"""Wide-ResNet WRN-28-4 on CIFAR-10 with Adam (ref ``keras_model_zoo/wresnet.py:37-82,159``):
pre-activation blocks, widths (16, 64, 128, 256), 4 blocks per group.  Self-contained
optimizer ⇒ only ``sync_type='avg'`` (``:152-153``); params = all trainable weights
(``:257-263``).  The model of the GOSGD benchmark config (BASELINE.json)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..torch_base import TorchModelBase

n_epochs = 200
batch_size = 128
file_batch_size = 128
learning_rate = 1e-3
depth, widen = 28, 4


class PreActBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.bn1, self.c1 = nn.BatchNorm2d(cin), nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn2, self.c2 = nn.BatchNorm2d(cout), nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.short = None if (stride == 1 and cin == cout) else nn.Conv2d(cin, cout, 1, stride, bias=False)

    def forward(self, x):
        o = torch.relu(self.bn1(x))
        s = x if self.short is None else self.short(o)
        o = self.c1(o)
        o = self.c2(torch.relu(self.bn2(o)))
        return o + s


class WRN(nn.Module):
    def __init__(self, depth=28, k=4, n_class=10):
        super().__init__()
        n = (depth - 4) // 6
        w = [16, 16 * k, 32 * k, 64 * k]
        layers = [nn.Conv2d(3, w[0], 3, 1, 1, bias=False)]
        cin = w[0]
        for g, (cout, stride) in enumerate(zip(w[1:], (1, 2, 2))):
            for j in range(n):
                layers.append(PreActBlock(cin, cout, stride if j == 0 else 1))
                cin = cout
        self.body = nn.Sequential(*layers)
        self.bn = nn.BatchNorm2d(cin)
        self.fc = nn.Linear(cin, n_class)

    def forward(self, x):
        return self.fc(torch.relu(self.bn(self.body(x))).mean((2, 3)))


class Wide_ResNet(TorchModelBase):
    n_epochs, batch_size, file_batch_size, learning_rate = n_epochs, batch_size, file_batch_size, learning_rate
    weight_decay, momentum = 0.0, 0.9
    lr_policy = "step"
    lr_step = [60, 120, 160]
    lr_gamma = 0.2
    input_width = input_height = 32

    def __init__(self, config):
        super().__init__(config)
        self.name = "Wide_ResNet"
        for k in ("batch_size", "file_batch_size", "n_epochs", "learning_rate"):
            if k in config:
                setattr(self, k, config[k])
        self.base_lr = self.learning_rate
        torch.manual_seed(23455)
        from ..data.cifar10 import Cifar10_data
        self.data = Cifar10_data(verbose=False, **config.get("data_kwargs", {}))
        self.channels = 3
        self.setup_data_parallel(self.data)
        self._mean = torch.as_tensor(self.data.rawdata[4]).to(self.device)
        net = WRN(config.get("depth", depth), config.get("widen", widen), self.data.n_class)
        self.finalize_torch(net, (self.batch_size, 32, 32, 3))

    def forward(self, x):
        x = (x.float() - self._mean) / 64.0
        return super().forward(x.to(self.act_dtype))

    def make_torch_optimizer(self, params):
        if self.config.get("optimizer", "adam") == "sgd":
            return None
        return torch.optim.Adam(params, lr=self.learning_rate)
