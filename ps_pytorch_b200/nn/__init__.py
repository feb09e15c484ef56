from .backend import set_backend, get_backend, backend, use_sm100  # noqa: F401
from .layers import (Conv2d, Linear, BatchNorm2d, ReLU, MaxPool2d, AvgPool2d, Dropout,  # noqa: F401
                     CrossEntropyLoss)
