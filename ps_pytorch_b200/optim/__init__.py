from .sgd import SGD  # noqa: F401
from .adam import Adam  # noqa: F401


def build_optimizer(name, params, lr, momentum=0.0, weight_decay=0.0, nesterov=False, dampening=0.0, amsgrad=False):
    if name == "sgd":
        return SGD(params, lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
    if name == "adam":
        return Adam(params, lr=lr, weight_decay=weight_decay, amsgrad=amsgrad)
    raise ValueError(name)
