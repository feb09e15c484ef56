"""Config parsing, launcher failure detection, data pipeline (native vs python), codec fuzz, reference-API objects."""
import os
import subprocess
import sys
import argparse

import numpy as np
import pytest
import torch

from ps_pytorch_b200.config import add_fit_args, add_eval_args, add_single_machine_args, str2bool, num_classes_for, input_shape_for

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flag_schema_matches_reference_defaults():
    a = add_fit_args(argparse.ArgumentParser(), [])
    assert (a.batch_size, a.test_batch_size, a.epochs, a.max_steps, a.lr, a.momentum, a.seed, a.log_interval) == (128, 500, 100, 10000, 0.01, 0.5, 1, 10)
    assert (a.network, a.mode, a.kill_threshold, a.dataset, a.comm_type, a.num_aggregate, a.eval_freq, a.train_dir, a.compress_grad) == \
        ("LeNet", "normal", 7.0, "MNIST", "Bcast", 5, 50, "output/models/", "compress")
    assert a.enable_gpu is False
    # the reference's launch line uses "--enable-gpu=" for False; any explicit true string is True
    assert add_fit_args(argparse.ArgumentParser(), ["--enable-gpu="]).enable_gpu is False
    assert add_fit_args(argparse.ArgumentParser(), ["--enable-gpu=true"]).enable_gpu is True
    e = add_eval_args(argparse.ArgumentParser(), [])
    assert (e.eval_batch_size, e.eval_freq, e.model_dir, e.dataset, e.network) == (10000, 50, "output/models/", "MNIST", "LeNet")
    s = add_single_machine_args(argparse.ArgumentParser(), [])
    assert s.test_batch_size == 1000
    assert str2bool("Yes") and not str2bool("none")
    with pytest.raises(argparse.ArgumentTypeError):
        str2bool("maybe")
    assert num_classes_for("Cifar100") == 100 and input_shape_for("MNIST") == (1, 28, 28)


def test_launcher_terminates_peers_when_a_rank_dies(tmp_path):
    script = tmp_path / "die.py"
    script.write_text("import os, sys, time\nif os.environ['RANK'] == '1':\n    sys.exit(3)\ntime.sleep(60)\n")
    r = subprocess.run([sys.executable, "-m", "ps_pytorch_b200.launch", "-n", "3", str(script)], cwd=ROOT,
                       env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=50)
    assert r.returncode == 3 and "rank 1 exited with 3" in r.stderr      # well under the 60 s the survivors would sleep


def test_batch_loader_native_equals_python_and_next_batch():
    from ps_pytorch_b200.utils.data import BatchLoader, synthetic_dataset
    ds = synthetic_dataset("Cifar10", 512)
    a = BatchLoader(ds, 64, shuffle=True, augment=True, seed=11, prefetch=0)
    b = BatchLoader(ds, 64, shuffle=True, augment=True, seed=11, prefetch=0)
    b._native = None
    if a._native is None:
        pytest.skip("host extension not built")
    for (xa, ya), (xb, yb) in zip(a, b):
        assert torch.equal(ya, yb) and torch.allclose(xa, xb, atol=1e-5)
    c = BatchLoader(ds, 100, shuffle=False, augment=False, prefetch=2)
    n = sum(x.size(0) for x, _ in c)
    assert n == 512 and ds.epochs_completed >= 1
    x, y = c.next_batch()
    assert x.shape == (100, 3, 32, 32) and y.shape == (100,)
    x2, _ = c.next_batch(batch_size=50)
    assert x2.shape[0] == 50


def test_codec_fuzz():
    from ps_pytorch_b200.utils import codec
    rng = np.random.RandomState(0)
    for n in (0, 1, 7, 13, 4096, 70001):
        for kind in ("rand", "zeros", "const", "sparse"):
            a = {"rand": rng.randn(n), "zeros": np.zeros(n), "const": np.full(n, 3.25), "sparse": rng.randn(n) * (rng.rand(n) < 0.05)}[kind].astype(np.float32)
            t = torch.from_numpy(a)
            out = torch.empty_like(t)
            codec.decompress_into(codec.compress(t), out)
            assert torch.equal(t, out), (n, kind)


def test_gradient_accumulator_and_model_buffer_api():
    from ps_pytorch_b200.models import build_model
    from ps_pytorch_b200.parallel import GradientAccumulator, ModelBuffer, ParamLayout
    m = build_model("LeNet", 10)
    lay = ParamLayout.from_module(m, channels_last=False)
    acc = GradientAccumulator(lay, num_worker=3, mode='None')
    assert len(acc.gradient_aggregator) == 8 and len(acc.gradient_aggregator[0]) == 3
    assert acc.gradient_aggregator[2][1].shape == m.conv2.weight.shape and acc.model_index_range == list(range(8))
    acc.gradient_aggregator[0][0].fill_(1.0)
    acc.gradient_aggregate_counter[0] = 2
    acc.meset_everything()
    assert acc.gradient_aggregate_counter == [0] * 8 and float(acc.slab.abs().sum()) == 0.0
    mb = ModelBuffer(m)
    assert mb.layer_cur_step == [0] * 8 and mb.layer_shape[0] == (20, 1, 5, 5)


def test_master_reference_api_aggregate_and_update():
    """aggregate_gradient / _model_update / meset_grad_buffer of the reference PS (sync_replicas_master_nn.py:204-262)."""
    from ps_pytorch_b200.parallel.ps import SyncReplicasMaster_NN
    from ps_pytorch_b200.parallel.transport import LocalTransport
    from ps_pytorch_b200.parallel.arena import ParamLayout
    from ps_pytorch_b200.models import build_model
    torch.manual_seed(0)
    net = build_model("LeNet", 10)
    comm = LocalTransport(ParamLayout.from_module(net, channels_last=False))
    ps = SyncReplicasMaster_NN(comm=comm, learning_rate=0.1, momentum=0.0, network="LeNet", kill_threshold=2, max_steps=1)
    ps.build_model(num_classes=10, network=net)
    before = [p.detach().clone() for p in ps.network.parameters()]
    for _ in range(2):
        for i, p in enumerate(ps.network.parameters()):
            ps.aggregate_gradient(np.ones(tuple(p.shape), dtype=np.float32), i)
    ps._model_update()                       # sum of 2 / K=2 = 1  ->  p -= 0.1
    for b, p in zip(before, ps.network.parameters()):
        assert torch.allclose(p, b - 0.1, atol=1e-6)
    ps.meset_grad_buffer()
    assert float(ps._grad_aggregate_buffer.flat.abs().sum()) == 0.0


def test_compression_facade_names():
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import numpy as np, compression as c; a=np.arange(12,dtype=np.float32).reshape(3,4); "
            "assert (c.g_decompress(c.g_compress(a))==a).all() and (c.w_decompress(c.w_compress(a))==a).all(); "
            "assert c._trim_msg(b'abc'+bytes([0x29])*32+b'junk')==b'abc' and c._trim_msg(b'abc')==b'abc'; print('ok')")
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.join(root, "src"), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-1500:]


def test_prometheus_metrics_endpoint():
    """--metrics-port: the worker log fields are scrapeable gauges (rank r on port+r)."""
    import socket, urllib.request
    from ps_pytorch_b200.utils.obs import MetricsExporter
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    m = MetricsExporter(port - 3, 3)
    assert m.on
    m.update(step=7, loss=0.25, dur=0.5, ips=2048.0, fwd=0.1, bwd=0.2, comm=0.05, msg_mb=1.5)
    m.update(step=8, loss=0.20, dur=0.5, ips=2048.0, fwd=0.1, bwd=0.2, comm=0.05, msg_mb=1.5)
    body = urllib.request.urlopen("http://127.0.0.1:%d/metrics" % port, timeout=10).read().decode()
    assert 'psb200_step{rank="3"} 8.0' in body and 'psb200_loss{rank="3"} 0.2' in body
    assert 'psb200_pushed_megabytes_total{rank="3"} 3.0' in body
    assert not MetricsExporter(0, 0).on


def test_pinned_pool_never_overwrites_a_held_or_in_flight_batch(monkeypatch):
    """the loader's pinned buffers are a pool with completion tracking.  A consumer that holds the current batch and
    the look-ahead, and whose H2D copies complete late, must never see a buffer rewritten under it (CPU model: pin_memory patched out,
    the 'event' is a fake whose synchronize() marks the copy as done)."""
    import torch
    from ps_pytorch_b200.utils import data as D
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    ds = D.synthetic_dataset("Cifar10", 16 * 23 + 5)
    ld = D.BatchLoader(ds, 16, shuffle=True, augment=True, seed=3, pin=False, prefetch=2)
    ld.pin = True                                   # no CUDA here: exercise the pool logic itself

    class FakeEvent:
        def __init__(self, key, log):
            self.key, self.log = key, log

        def synchronize(self):
            self.log.add(self.key)

    synced = set()
    for epoch in range(2):
        held = []                                   # (tensor, snapshot, batch id)
        it = iter(ld)
        nxt = next(it, None)
        i = 0
        while nxt is not None:
            cur = nxt
            cur[0]._psb_evt = FakeEvent((epoch, i), synced)      # "H2D copy enqueued"
            held.append((cur[0], cur[0].clone(), (epoch, i)))
            nxt = next(it, None)                    # look-ahead, as DistributedWorker.train does
            for t, snap, key in held:
                if key not in synced:               # copy not complete yet: the buffer must be untouched
                    assert torch.equal(t, snap), ("buffer of batch %s rewritten before its copy completed" % (key,))
            held = [h for h in held if h[2] not in synced]
            i += 1
        assert i == len(ld)
    assert len(ld._pool) == ld.hold + 2 + 2
