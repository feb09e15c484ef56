"""All-reduce DP wrapper on the gloo tier: averaged gradients equal the single-process gradient of the concatenated batch."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from ps_pytorch_b200.models import build_model
from ps_pytorch_b200.parallel.ddp import DistributedDataParallel
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
torch.manual_seed(100 + r)                      # different init per rank: the wrapper must broadcast rank 0's
ddp = DistributedDataParallel(build_model("LeNet", 10), "cpu")
g = torch.Generator().manual_seed(5 + r)
x, y = torch.randn(8, 1, 28, 28, generator=g), torch.randint(0, 10, (8,), generator=g)
ddp.zero_grad()
torch.nn.functional.cross_entropy(ddp(x), y).backward()
ddp.sync_gradients()
if r == 0:
    torch.save({"w": [p.detach().clone() for p in ddp.module.parameters()], "g": [p.grad.clone() for p in ddp.module.parameters()]}, sys.argv[1])
dist.barrier()
'''


def test_ddp_gloo(tmp_path):
    script = tmp_path / "ddp_job.py"
    script.write_text(SCRIPT % ROOT)
    out = str(tmp_path / "o.pt")
    r = subprocess.run([sys.executable, "-m", "ps_pytorch_b200.launch", "-n", "2", "--timeout", "100", str(script), out],
                       cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = torch.load(out)
    from ps_pytorch_b200.models import build_model
    torch.manual_seed(100)
    m = build_model("LeNet", 10)
    for a, b in zip(res["w"], m.parameters()):
        assert torch.equal(a, b)
    tot = 0
    for rk in range(2):
        g = torch.Generator().manual_seed(5 + rk)
        x, y = torch.randn(8, 1, 28, 28, generator=g), torch.randint(0, 10, (8,), generator=g)
        tot = tot + torch.nn.functional.cross_entropy(m(x), y) / 2
    tot.backward()
    for a, p in zip(res["g"], m.parameters()):
        assert torch.allclose(a, p.grad, atol=1e-6)


def test_ddp_cli_mode_and_reference_facades(tmp_path):
    """``--parallelism ddp`` through the reference-shaped CLI: 2 gloo ranks train LeNet, the loss falls, a checkpoint the evaluator can read
    appears; and the reference's module paths for the vendored DDP / autograd backward resolve to the live implementations."""
    import subprocess
    d = str(tmp_path) + "/"
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", "2", "--timeout", "200", os.path.join(ROOT, "src", "distributed_nn.py"),
           "--parallelism", "ddp", "--network", "LeNet", "--dataset", "MNIST", "--batch-size", "32", "--max-steps", "20", "--lr", "0.05",
           "--momentum", "0.9", "--eval-freq", "20", "--train-dir", d, "--log-interval", "1", "--enable-gpu="]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout + r.stderr
    from ps_pytorch_b200.utils.obs import parse_worker_line
    recs = [x for x in (parse_worker_line(l) for l in out.splitlines()) if x and x["rank"] == 0]
    assert len(recs) == 20 and recs[-1]["loss"] < recs[0]["loss"]
    assert os.path.isfile(d + "model_step_20")
    sys.path.insert(0, os.path.join(ROOT, "src"))
    try:
        from data_parallel_dist import DistributedDataParallel as D1
        from distributed_functions import backward
    finally:
        sys.path.pop(0)
    from ps_pytorch_b200.parallel.ddp import DistributedDataParallel as D2
    assert D1 is D2
    lin = torch.nn.Linear(3, 2)
    seen = []
    backward(lin(torch.ones(1, 3)).sum(), on_grad=lambda p, g: seen.append(tuple(g.shape)))
    assert sorted(seen) == [(2,), (2, 3)] and lin.weight.grad is not None
