import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ps_pytorch_b200.models.resnet import ResNetSplit, ResNetSplit18, ResNetSplit34, ResNetSplit50  # noqa: F401,E402
