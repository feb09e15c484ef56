"""Model zoo + factory.  ``build_model(name, num_classes)`` keeps the reference's names
(``/root/reference/src/util.py:8-19``: LeNet, ResNet18, ResNet34, ResNet50, VGG11→vgg11_bn) and widens
them to everything the reference README claims (README.md:124) plus the BASELINE.json configs
(VGG-16, ImageNet-shaped ResNet-50).  The reference's ``ResNet34`` factory bug (missing argument)
and fixed-10-class ``ResNet50`` are fixed; an unknown name raises instead of returning ``None``."""
from __future__ import annotations

from .lenet import LeNet, LeNetSplit
from .resnet import (ResNet, BasicBlock, Bottleneck, ResNet18, ResNet34, ResNet50, ResNet101, ResNet152,
                     ResNetSplit, ResNetSplit18, ResNetSplit34, ResNetSplit50)
from .vgg import VGG, vgg11, vgg11_bn, vgg13, vgg13_bn, vgg16, vgg16_bn, vgg19, vgg19_bn
from .layerwise import LayerwiseBackwardMixin, StragglerKilled

_FACTORY = {
    "LeNet": lambda nc: LeNet(nc),
    "LeNetSplit": lambda nc: LeNetSplit(nc),
    "ResNet18": lambda nc: ResNet18(nc),
    "ResNet34": lambda nc: ResNet34(nc),
    "ResNet50": lambda nc: ResNet50(nc),
    "ResNet101": lambda nc: ResNet101(nc),
    "ResNet152": lambda nc: ResNet152(nc),
    "ResNet": lambda nc: ResNetSplit18(num_classes=nc),        # single_machine.py's "--network ResNet"
    "ResNetSplit18": lambda nc: ResNetSplit18(num_classes=nc),
    "ResNetSplit34": lambda nc: ResNetSplit34(num_classes=nc),
    "ResNetSplit50": lambda nc: ResNetSplit50(num_classes=nc),
    "ResNet18ImageNet": lambda nc: ResNet18(nc, imagenet=True),
    "ResNet50ImageNet": lambda nc: ResNet50(nc, imagenet=True),
    "VGG11": lambda nc: vgg11_bn(nc),                          # reference maps VGG11 → vgg11_bn
    "VGG11_bn": lambda nc: vgg11_bn(nc), "VGG11_plain": lambda nc: vgg11(nc),
    "VGG13": lambda nc: vgg13(nc), "VGG13_bn": lambda nc: vgg13_bn(nc),
    "VGG16": lambda nc: vgg16(nc), "VGG16_bn": lambda nc: vgg16_bn(nc),
    "VGG19": lambda nc: vgg19(nc), "VGG19_bn": lambda nc: vgg19_bn(nc),
}


def build_model(model_name: str, num_classes: int = 10):
    try:
        return _FACTORY[model_name](num_classes)
    except KeyError:
        raise ValueError("unknown network %r; choose from %s" % (model_name, sorted(_FACTORY))) from None


def model_names():
    return sorted(_FACTORY)
