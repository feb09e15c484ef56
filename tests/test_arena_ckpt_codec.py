import os

import numpy as np
import torch

from ps_pytorch_b200.models import build_model
from ps_pytorch_b200.parallel.arena import ALIGN, FlatBuffer, ParamLayout
from ps_pytorch_b200.utils import checkpoint as ckpt, codec
from ps_pytorch_b200.utils.obs import format_worker_line, parse_worker_line


def test_layout_alignment_and_views():
    m = build_model("ResNet18", 10)
    lay = ParamLayout.from_module(m, channels_last=True)
    assert all(o % ALIGN == 0 for o in lay.offsets) and lay.total >= lay.num_params
    ref = [p.detach().clone() for p in m.parameters()]
    fb = FlatBuffer(lay, torch.float32, "cpu")
    fb.bind_params(m)
    for p, r in zip(m.parameters(), ref):
        assert torch.equal(p, r) and p.data_ptr() >= fb.flat.data_ptr()
    w = m.conv1.weight                                # [64,3,3,3] logical, KRSC physical
    assert w.shape == (64, 3, 3, 3) and w.stride() == (27, 1, 9, 3)
    bk = lay.buckets(4)
    assert bk[0][0] == 0 and bk[-1][1] == lay.total and sum(len(b[2]) for b in bk) == 62
    for (lo, hi, idx), nxt in zip(bk, bk[1:]):
        assert hi == nxt[0]


def test_grad_binding_accumulates_in_arena():
    m = build_model("LeNet", 10)
    lay = ParamLayout.from_module(m, channels_last=False)
    g = FlatBuffer(lay, torch.float32, "cpu")
    g.bind_grads(m)
    m(torch.randn(2, 1, 28, 28)).sum().backward()
    assert g.flat.abs().sum() > 0
    assert m.fc2.bias.grad.data_ptr() == g.views[-1].data_ptr()


def test_checkpoint_roundtrip_and_reference_layout(tmp_path):
    m = build_model("ResNet18", 10)
    d = str(tmp_path) + "/"
    p = ckpt.save_model(m, d, 50)
    assert p == d + "model_step_50" and os.path.isfile(p)
    sd = torch.load(p)
    assert all(v.dtype in (torch.float32, torch.int64) and v.is_contiguous() for v in sd.values())
    m2 = build_model("ResNet18", 10)
    ckpt.load_model_into(m2, p)
    for a, b in zip(m.state_dict().values(), m2.state_dict().values()):
        assert torch.equal(a, b)
    # stock torch.nn module with the reference's structure can load it too (evaluator compatibility)
    import torch.nn as nn
    ref = nn.Sequential()
    ref.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
    ref.load_state_dict({"conv1.weight": sd["conv1.weight"]})


def test_codec_lossless_and_reference_api():
    g = torch.randn(1000, 37)
    blob = codec.compress(g)
    out = torch.empty_like(g)
    codec.decompress_into(blob, out)
    assert torch.equal(g, out)
    z = torch.zeros(100000)
    assert codec.compress(z).numel() < z.numel() * 4 / 20
    a = np.random.randn(5, 3, 2).astype(np.float32)
    assert np.array_equal(codec.g_decompress(codec.g_compress(a)), a)


def test_log_line_roundtrip():
    line = format_worker_line(3, 17, 1, 2048, 50000, 1.2345, 0.5, 0.1, 0.2, 0.15, 0.05, 22.3, 2048.0)
    d = parse_worker_line("INFO:x:" + line)
    assert d["rank"] == 3 and d["step"] == 17 and abs(d["loss"] - 1.2345) < 1e-9 and abs(d["msg_mb"] - 22.3) < 1e-9
    assert line.startswith("Worker: 3, Step: 17, Epoch: 1 [2048/50000 (4%)], Loss: 1.2345, Time Cost: 0.5000, FetchWeight: 0.1000")
