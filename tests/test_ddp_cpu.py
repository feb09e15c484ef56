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
