"""VGG 11/13/16/19 (±BN) with the CIFAR classifier of the reference
(``/root/reference/src/model_ops/vgg.py:15-108``): features → Dropout, Linear(512,512), ReLU,
Dropout, Linear(512,512), ReLU, Linear(512,nc); conv init N(0, sqrt(2/(k·k·Cout))), bias 0.
``state_dict`` keys (``features.N`` / ``classifier.N``) are those of the reference."""
from __future__ import annotations

import math

import torch.nn as tnn

from .. import nn

cfg = {
    'A': [64, 'M', 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'B': [64, 64, 'M', 128, 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'D': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],
    'E': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M'],
}


class _Features(tnn.Sequential):
    """Runs conv→BN→ReLU triples through the fused BN(+ReLU) call while keeping Sequential indices."""

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(m, (nn.BatchNorm2d, nn.Conv2d)) and isinstance(nxt, tnn.ReLU):
                x = m(x, relu=True)          # BN+ReLU in one pass / bias+ReLU in the conv epilogue
                i += 2
            elif isinstance(m, nn.Conv2d) and isinstance(nxt, nn.BatchNorm2d):
                x = m(x, bn=nxt)             # the BN's batch statistics are reduced in the convolution epilogue
                i += 1
            else:
                x = m(x)
                i += 1
        return x


class _Classifier(tnn.Sequential):
    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(m, nn.Linear) and isinstance(nxt, tnn.ReLU):
                x = m(x, relu=True)
                i += 2
            else:
                x = m(x)
                i += 1
        return x


class VGG(tnn.Module):
    def __init__(self, features, num_classes=10):
        super().__init__()
        self.features = features
        self.classifier = _Classifier(
            nn.Dropout(), nn.Linear(512, 512), nn.ReLU(True),
            nn.Dropout(), nn.Linear(512, 512), nn.ReLU(True),
            nn.Linear(512, num_classes))
        for m in self.modules():
            if isinstance(m, tnn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
                m.bias.data.zero_()

    def forward(self, x):
        x = self.features(x)
        return self.classifier(x.reshape(x.size(0), -1))


def make_layers(c, batch_norm=False):
    layers, in_channels = [], 3
    for v in c:
        if v == 'M':
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            conv = nn.Conv2d(in_channels, v, kernel_size=3, padding=1)
            layers += [conv, nn.BatchNorm2d(v), nn.ReLU(inplace=True)] if batch_norm else [conv, nn.ReLU(inplace=True)]
            in_channels = v
    return _Features(*layers)


def _mk(key, bn):
    def ctor(num_classes=10):
        return VGG(make_layers(cfg[key], batch_norm=bn), num_classes=num_classes)
    return ctor


vgg11, vgg11_bn = _mk('A', False), _mk('A', True)
vgg13, vgg13_bn = _mk('B', False), _mk('B', True)
vgg16, vgg16_bn = _mk('D', False), _mk('D', True)
vgg19, vgg19_bn = _mk('E', False), _mk('E', True)
