"""Optimizer math: ``step(grads)`` with hand-fed gradients vs torch.optim (SURVEY.md §2.4 semantics)."""
import numpy as np
import pytest
import torch

from ps_pytorch_b200.optim import SGD, Adam
from ps_pytorch_b200.parallel.arena import FlatBuffer, ParamLayout


def _mk():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))


@pytest.mark.parametrize("kw", [dict(momentum=0.0), dict(momentum=0.9), dict(momentum=0.9, nesterov=True),
                                dict(momentum=0.9, weight_decay=1e-2), dict(momentum=0.5, dampening=0.3)])
def test_sgd_matches_torch(kw):
    a, b = _mk(), _mk()
    oa = SGD(a.parameters(), lr=0.1, **kw)
    ob = torch.optim.SGD(b.parameters(), lr=0.1, **kw)
    for it in range(4):
        grads = [np.random.RandomState(it * 10 + i).randn(*p.shape).astype(np.float32) for i, p in enumerate(a.parameters())]
        oa.step(grads=grads)                     # numpy arrays, like the reference PS
        for p, g in zip(b.parameters(), grads):
            p.grad = torch.from_numpy(g)
        ob.step()
        for pa, pb in zip(a.parameters(), b.parameters()):
            # torch applies dampening on the first step too? no: first step buf = grad for both
            assert torch.allclose(pa, pb, atol=1e-6), (it, kw)


@pytest.mark.parametrize("kw", [dict(), dict(weight_decay=1e-2), dict(amsgrad=True)])
def test_adam_matches_torch(kw):
    a, b = _mk(), _mk()
    oa, ob = Adam(a.parameters(), lr=1e-2, **kw), torch.optim.Adam(b.parameters(), lr=1e-2, **kw)
    for it in range(5):
        grads = [torch.randn(p.shape, generator=torch.Generator().manual_seed(it * 10 + i)) for i, p in enumerate(a.parameters())]
        oa.step(grads=grads)
        for p, g in zip(b.parameters(), grads):
            p.grad = g.clone()
        ob.step()
        for pa, pb in zip(a.parameters(), b.parameters()):
            assert torch.allclose(pa, pb, atol=2e-6), (it, kw)


@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_flat_path_equals_list_path(opt):
    a, b = _mk(), _mk()
    layout = ParamLayout.from_module(a, channels_last=False, chunk=256)
    master = FlatBuffer(layout, torch.float32, "cpu")
    master.bind_params(a)
    mk = (lambda ps: SGD(ps, lr=0.05, momentum=0.9, weight_decay=1e-3)) if opt == "sgd" else (lambda ps: Adam(ps, lr=1e-2, amsgrad=True))
    oa, ob = mk(a.parameters()), mk(b.parameters())
    oa.attach_flat(master)
    gb = FlatBuffer(layout, torch.float32, "cpu")
    for it in range(3):
        grads = [torch.randn(p.shape, generator=torch.Generator().manual_seed(it * 7 + i)) for i, p in enumerate(b.parameters())]
        gb.load_list(grads)
        oa.step(grads=gb.flat, scale=0.5)
        ob.step(grads=[g * 0.5 for g in grads])
        for pa, pb in zip(a.parameters(), b.parameters()):
            assert torch.allclose(pa, pb, atol=1e-6)


def test_ctor_validation():
    with pytest.raises(ValueError):
        SGD(_mk().parameters(), lr=-1)
    with pytest.raises(ValueError):
        SGD(_mk().parameters(), lr=0.1, nesterov=True)
