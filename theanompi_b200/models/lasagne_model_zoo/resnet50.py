"""ResNet-50 (ref ``theanompi/models/lasagne_model_zoo/resnet50.py``): bottleneck blocks
[3,4,6,3], the last BatchNorm gamma of every residual branch initialised to 0 (``:62-65``),
batch 32 (the published table uses 64), lr 0.1·b/256, μ 0.9, wd 1e-4, momentum-SGD through
the framework's ``pre_model_iter_fn`` path.  BN gamma/beta are updated locally and never
exchanged (``opt.py:207-226``, ``exchanger.py:35-43``)."""
from __future__ import annotations

import torch.nn as nn

from ..torch_base import TorchModelBase

n_epochs = 90
momentum = 0.9
weight_decay = 1e-4
batch_size = 32
file_batch_size = 128
learning_rate = 0.1 * batch_size / 256.0
lr_policy = "step"
lr_step = [30, 60, 80]
input_width = input_height = 224


class Bottleneck(nn.Module):
    def __init__(self, cin, mid, stride):
        super().__init__()
        cout = mid * 4
        self.a = nn.Sequential(nn.Conv2d(cin, mid, 1, bias=False), nn.BatchNorm2d(mid), nn.ReLU(inplace=True),
                               nn.Conv2d(mid, mid, 3, stride, 1, bias=False), nn.BatchNorm2d(mid), nn.ReLU(inplace=True),
                               nn.Conv2d(mid, cout, 1, bias=False), nn.BatchNorm2d(cout))
        nn.init.zeros_(self.a[-1].weight)                  # 2c branch gamma = 0 (ref :62-65)
        self.proj = None
        if stride != 1 or cin != cout:
            self.proj = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu(self.a(x) + (x if self.proj is None else self.proj(x)))


class ResNet50Net(nn.Module):
    def __init__(self, n_class=1000, blocks=(3, 4, 6, 3)):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                  nn.MaxPool2d(3, 2, 1))
        layers, cin = [], 64
        for i, n in enumerate(blocks):
            mid = 64 * 2 ** i
            for j in range(n):
                layers.append(Bottleneck(cin, mid, 2 if (j == 0 and i > 0) else 1))
                cin = mid * 4
        self.body = nn.Sequential(*layers)
        self.head = nn.Linear(cin, n_class)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        return self.head(self.body(self.stem(x)).mean((2, 3)))


class ResNet50(TorchModelBase):
    n_epochs, momentum, weight_decay = n_epochs, momentum, weight_decay
    batch_size, file_batch_size, learning_rate = batch_size, file_batch_size, learning_rate
    lr_policy, lr_step = lr_policy, lr_step
    input_width, input_height = input_width, input_height
    blocks = (3, 4, 6, 3)

    def __init__(self, config):
        super().__init__(config)
        self.name = "ResNet50"
        for k in ("batch_size", "file_batch_size", "n_epochs"):
            if k in config:
                setattr(self, k, config[k])
        import torch
        torch.manual_seed(23455)
        from ..data.imagenet import ImageNet_data
        dk = dict(config.get("data_kwargs", {}))
        if "n_class" in config:
            dk.setdefault("n_class", config["n_class"])
        self.data = ImageNet_data(verbose=False, file_batch_size=self.file_batch_size, **dk)
        self.channels = self.data.channels
        self.setup_data_parallel(self.data)
        net = ResNet50Net(config.get("n_class", self.data.n_class), config.get("blocks", self.blocks))
        self.finalize_torch(net, (self.batch_size, self.input_height, self.input_width, self.channels))
        if self.data.para_load and not self.no_paraload:
            self.data.spawn_load()
            self.data.para_load_init(self.device, self.input_width, self.input_height, self.rand_crop,
                                     self.batch_crop_mirror, out_dtype=self.act_dtype)
