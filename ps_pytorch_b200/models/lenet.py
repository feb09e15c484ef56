"""LeNet for 1x28x28 inputs.  Architecture and ``state_dict`` keys follow the reference
(``/root/reference/src/model_ops/lenet.py:16-37``): conv(1→20,k5) → maxpool2 → relu →
conv(20→50,k5) → maxpool2 → relu → fc(800→500) → fc(500→10), no activation between the fcs.
"""
from __future__ import annotations

import torch.nn as tnn

from .. import nn
from .layerwise import LayerwiseBackwardMixin


class LeNet(tnn.Module):
    def __init__(self, num_classes: int = 10):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 20, 5, 1)
        self.conv2 = nn.Conv2d(20, 50, 5, 1)
        self.fc1 = nn.Linear(4 * 4 * 50, 500)
        self.fc2 = nn.Linear(500, num_classes)
        self.pool = nn.MaxPool2d(2, 2)

    def forward(self, x):
        x = self.pool(self.conv1(x), relu=True)
        x = self.pool(self.conv2(x), relu=True)
        # flatten in NCHW (c,h,w) order whatever the memory format, as the reference's view(-1, 800) does
        x = x.contiguous().view(x.size(0), -1) if x.is_contiguous() else x.reshape(x.size(0), -1)
        return self.fc2(self.fc1(x))

    def name(self):
        return 'lenet'


class LeNetSplit(LayerwiseBackwardMixin, LeNet):
    """LeNet with the layer-by-layer backward API of the reference's ``LeNetSplit``
    (``/root/reference/src/model_ops/lenet.py:39-257``): gradients are handed to the
    communicator as soon as each layer's backward has produced them (last layer first),
    optionally polling the PS kill signal between layers."""

    def name(self):
        return 'lenet_split'
