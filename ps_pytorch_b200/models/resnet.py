"""ResNets.  The CIFAR variant (3x3 stem, no max-pool, 4x4 average pool) has the module
names, shapes and parameter order of the reference (``/root/reference/src/model_ops/resnet.py:14-113``)
so its checkpoints are interchangeable; ``imagenet=True`` adds the 7x7/2 stem + 3x3/2 max-pool +
global average pool needed by the ImageNet-shaped ResNet-50 config in BASELINE.json (the
reference model cannot take 224x224 inputs).

BN → (+shortcut) → ReLU is expressed through the fused ``nn.BatchNorm2d(x, residual, relu)`` call.
"""
from __future__ import annotations

import torch.nn as tnn
import torch.nn.functional as F

from .. import nn
from ..nn.backend import use_sm100
from ..ops import conv as _conv
from .layerwise import LayerwiseBackwardMixin


class _Shortcut(tnn.Sequential):
    """``shortcut.0`` conv + ``shortcut.1`` BN (empty Sequential = identity), as in the reference."""

    def forward(self, x):
        if len(self) == 0:
            return x
        return self[1](self[0](x, bn=self[1]))


class BasicBlock(tnn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.shortcut = _Shortcut()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = _Shortcut(
                nn.Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(self.expansion * planes))

    def forward(self, x):
        # The gradient arriving through the skip path is parked and added inside conv1's dgrad epilogue (ops/conv.py:stash_grad).  The stash node
        # is created AFTER conv2: autograd runs ready nodes newest-first, so the skip gradient is parked before conv1's backward looks for it.
        h = _conv.new_grad_holder(x) if use_sm100(x) else None
        out = self.bn1(self.conv1(x, grad_holder=h, bn=self.bn1), relu=True)
        out = self.conv2(out, bn=self.bn2)
        xs = _conv.stash_grad(x, h) if h is not None else x
        return self.bn2(out, residual=self.shortcut(xs), relu=True)


class Bottleneck(tnn.Module):
    expansion = 4

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, self.expansion * planes, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(self.expansion * planes)
        self.shortcut = _Shortcut()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = _Shortcut(
                nn.Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(self.expansion * planes))

    def forward(self, x):
        h = _conv.new_grad_holder(x) if use_sm100(x) else None
        out = self.bn1(self.conv1(x, grad_holder=h, bn=self.bn1), relu=True)
        out = self.bn2(self.conv2(out, bn=self.bn2), relu=True)
        out = self.conv3(out, bn=self.bn3)
        xs = _conv.stash_grad(x, h) if h is not None else x          # created last: runs first in backward (see BasicBlock)
        return self.bn3(out, residual=self.shortcut(xs), relu=True)


class ResNet(tnn.Module):
    def __init__(self, block, num_blocks, num_classes=10, imagenet=False):
        super().__init__()
        self.in_planes = 64
        self.imagenet = imagenet
        if imagenet:
            self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        else:
            self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = self._make_layer(block, 64, num_blocks[0], stride=1)
        self.layer2 = self._make_layer(block, 128, num_blocks[1], stride=2)
        self.layer3 = self._make_layer(block, 256, num_blocks[2], stride=2)
        self.layer4 = self._make_layer(block, 512, num_blocks[3], stride=2)
        self.linear = nn.Linear(512 * block.expansion, num_classes)
        self.pool = nn.AvgPool2d(4)  # parameter-free: does not change the state_dict
        self.stem_pool = nn.MaxPool2d(3, 2, 1) if imagenet else None

    def _make_layer(self, block, planes, num_blocks, stride):
        layers = []
        for s in [stride] + [1] * (num_blocks - 1):
            layers.append(block(self.in_planes, planes, s))
            self.in_planes = planes * block.expansion
        return tnn.Sequential(*layers)

    def forward(self, x):
        out = self.bn1(self.conv1(x, bn=self.bn1), relu=True)
        if self.imagenet:
            out = self.stem_pool(out)
        out = self.layer4(self.layer3(self.layer2(self.layer1(out))))
        if self.imagenet:
            out = _global_mean(out)
        else:
            out = self.pool(out).reshape(out.size(0), -1)
        return self.linear(out)


def _global_mean(x):
    """Global average pool -> [N, C] (our kernel on the sm100 backend, ATen otherwise)."""
    if use_sm100(x) and x.dim() == 4 and x.size(1) % 8 == 0 and str(x.dtype) == "torch.bfloat16":
        from ..ops import elementwise as _ew
        return _ew.global_avg_pool(x).reshape(x.size(0), -1)
    return x.mean(dim=(2, 3))


def ResNet18(num_classes=10, imagenet=False):
    return ResNet(BasicBlock, [2, 2, 2, 2], num_classes, imagenet)


def ResNet34(num_classes=10, imagenet=False):
    return ResNet(BasicBlock, [3, 4, 6, 3], num_classes, imagenet)


def ResNet50(num_classes=10, imagenet=False):
    return ResNet(Bottleneck, [3, 4, 6, 3], num_classes, imagenet)


def ResNet101(num_classes=10, imagenet=False):
    return ResNet(Bottleneck, [3, 4, 23, 3], num_classes, imagenet)


def ResNet152(num_classes=10, imagenet=False):
    return ResNet(Bottleneck, [3, 8, 36, 3], num_classes, imagenet)


class ResNetSplit(LayerwiseBackwardMixin, ResNet):
    """ResNet with the reference's layer-wise backward API (``/root/reference/src/model_ops/resnet_split.py:142-749``):
    ``backward_normal`` / ``backward_signal_kill`` / ``backward_timeout_kill`` / ``backward_single``.
    Unlike the reference (only ``ResNetSplit18`` is constructible there) every depth works."""


def ResNetSplit18(kill_threshold=7.0, num_classes=10):
    m = ResNetSplit(BasicBlock, [2, 2, 2, 2], num_classes)
    m.kill_threshold = kill_threshold
    return m


def ResNetSplit34(kill_threshold=7.0, num_classes=10):
    m = ResNetSplit(BasicBlock, [3, 4, 6, 3], num_classes)
    m.kill_threshold = kill_threshold
    return m


def ResNetSplit50(kill_threshold=7.0, num_classes=10):
    m = ResNetSplit(Bottleneck, [3, 4, 6, 3], num_classes)
    m.kill_threshold = kill_threshold
    return m
