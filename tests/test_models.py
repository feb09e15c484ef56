"""Model factory: parameter-tensor counts / parameter counts from SURVEY.md §2.5 and forward shapes."""
import pytest
import torch

from ps_pytorch_b200.models import build_model, model_names

EXPECT = {"LeNet": (8, 431080), "ResNet18": (62, 11173962), "ResNet34": (110, None), "ResNet50": (161, 23520842),
          "ResNet101": (314, None), "ResNet152": (467, None), "VGG11": (38, 9756426), "VGG16": (32, 15245130),
          "VGG16_bn": (58, None), "VGG19_bn": (70, None)}


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_param_counts(name):
    m = build_model(name, 10)
    ps = list(m.parameters())
    nt, npar = EXPECT[name]
    assert len(ps) == nt
    if npar is not None:
        assert sum(p.numel() for p in ps) == npar


@pytest.mark.parametrize("name,shape,nc", [("LeNet", (2, 1, 28, 28), 10), ("ResNet18", (2, 3, 32, 32), 100),
                                           ("VGG11", (2, 3, 32, 32), 10), ("VGG16", (2, 3, 32, 32), 10),
                                           ("ResNet50ImageNet", (1, 3, 64, 64), 1000), ("ResNetSplit18", (2, 3, 32, 32), 10)])
def test_forward_shapes(name, shape, nc):
    m = build_model(name, nc).eval()
    with torch.no_grad():
        y = m(torch.randn(*shape))
    assert y.shape == (shape[0], nc)


def test_unknown_model_raises():
    with pytest.raises(ValueError):
        build_model("NoSuchNet", 10)
    assert "ResNet18" in model_names()


def test_state_dict_keys_match_reference_names():
    keys = set(build_model("ResNet18", 10).state_dict().keys())
    for k in ["conv1.weight", "bn1.running_mean", "layer1.0.conv1.weight", "layer2.0.shortcut.0.weight",
              "layer2.0.shortcut.1.running_var", "linear.bias"]:
        assert k in keys
    keys = set(build_model("VGG11", 10).state_dict().keys())
    assert "features.0.weight" in keys and "classifier.6.bias" in keys
    assert set(build_model("LeNet").state_dict().keys()) == {"conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
                                                              "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"}


def test_layerwise_backward_matches_autograd():
    torch.manual_seed(0)
    a, b = build_model("ResNetSplit18", 10), build_model("ResNet18", 10)
    b.load_state_dict(a.state_dict())
    x, y = torch.randn(4, 3, 32, 32), torch.randint(0, 10, (4,))
    la = torch.nn.functional.cross_entropy(a(x), y)
    g = torch.autograd.grad(la, a.output[-1])[0]
    pushed = []

    class Comm:
        def push_params(self, idx, step): pushed.append(list(idx))
        def poll_kill(self, step): return False
    a.backward_normal(g, Comm(), None, 3)
    torch.nn.functional.cross_entropy(b(x), y).backward()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa.grad, pb.grad, atol=1e-5, rtol=1e-4)
    flat = [i for s in pushed for i in s]
    assert sorted(flat) == list(range(62)) and flat[0] >= 60     # last layer first
    # kill: abandon immediately
    la = torch.nn.functional.cross_entropy(a(x), y)
    g = torch.autograd.grad(la, a.output[-1])[0]

    class Kill(Comm):
        def poll_kill(self, step): return True
    _, killed = a.backward_signal_kill(g, Kill(), None, 4)
    assert killed
