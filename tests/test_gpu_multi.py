"""Multi-GPU transport tests (2+ B200s): the NVLink tier must produce the update the gloo tier / a single-process
fp32 reference produces, in both topologies, with bf16 / fp8 wires, with and without multicast."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nprocs, out, extra, env_extra=None, timeout=300):
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", str(nprocs), "--timeout", str(timeout - 30),
           os.path.join(ROOT, "tests", "mp_job.py"), "--out", out, "--transport", "nvl", "--device", "cuda", "--dtype", "bf16"] + extra
    env = dict(os.environ, PYTHONPATH=ROOT)
    env.update(env_extra or {})
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return torch.load(out)


def _reference(nworkers, steps, bs, network):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_distributed_cpu import _reference as ref
    return ref(nworkers, steps, bs, network)


def _check(res, ref, init, tol_cos=0.97):
    # compare the *update* (w - w0): bf16 compute makes element-wise equality meaningless, direction must agree
    d_res = torch.cat([(a - i).flatten() for a, i in zip(res["w"], init)])
    d_ref = torch.cat([(b - i).flatten() for b, i in zip(ref, init)])
    cos = torch.nn.functional.cosine_similarity(d_res, d_ref, dim=0).item()
    rel = ((d_res - d_ref).norm() / d_ref.norm()).item()
    assert cos > tol_cos, (cos, rel)
    return cos, rel


def _init(network):
    from ps_pytorch_b200.models import build_model
    torch.manual_seed(7)
    return [p.detach().clone() for p in build_model(network, 10).parameters()]


@pytest.mark.parametrize("topology,wire,mc", [("ps", "bf16", "1"), ("ps", "fp8", "1"), ("ps", "bf16", "0"), ("sharded", "bf16", "1")])
def test_nvl_two_gpus_matches_reference(tmp_path, topology, wire, mc):
    n = min(torch.cuda.device_count(), 2)
    res = _run(n, str(tmp_path / "w.pt"), ["--steps", "3", "--network", "LeNet", "--bs", "64", "--topology", topology,
                                           "--grad-dtype", wire, "--colocate", "1", "--backend", "library"],
               env_extra={"PSB200_MULTICAST": mc})
    assert res["workers"] == n
    ref = _reference(n, 3, 64, "LeNet")
    _check(res, ref, _init("LeNet"), 0.97 if wire != "fp8" else 0.93)


def test_nvl_reference_topology_pure_ps(tmp_path):
    """--no-colocate: rank 0 only serves (the reference's topology)."""
    n = min(torch.cuda.device_count(), 2)
    res = _run(n, str(tmp_path / "w.pt"), ["--steps", "3", "--network", "LeNet", "--bs", "64", "--colocate", "0", "--backend", "library"])
    assert res["workers"] == n - 1
    _check(res, _reference(n - 1, 3, 64, "LeNet"), _init("LeNet"))


def test_nvl_resnet18_sm100_graph(tmp_path):
    n = min(torch.cuda.device_count(), 2)
    res = _run(n, str(tmp_path / "w.pt"), ["--steps", "2", "--network", "ResNet18", "--bs", "32", "--colocate", "1", "--graph", "1",
                                           "--backend", "sm100", "--overlap", "1"])
    assert res["workers"] == n and all(torch.isfinite(w).all() for w in res["w"])
    assert len(res["losses"]) == 2 and all(l == l for l in res["losses"])


@pytest.mark.parametrize("topology,mode,size,steps", [("ps", "push", "small", 300), ("ps", "pull", "small", 300), ("sharded", "push", "small", 300),
                                                      ("ps", "push", "resnet18", 60), ("sharded", "push", "resnet18", 60)])
def test_flag_protocol_litmus(topology, mode, size, steps):
    """Exact-arithmetic push/reduce/broadcast cycles with a per-element pattern, checked on the device every step without a host
    sync; the `resnet18` size runs the protocol on an 11.2 M-element arena (hundreds of CTAs per kernel)."""
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", str(n), "--timeout", "200", os.path.join(ROOT, "tests", "mp_litmus.py"), topology, str(steps), mode, size]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=260)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_flag_protocol_litmus_k_of_p_staggered():
    """K = P - 1 with workers that arrive in a different order every step: every update must be the mean of ONE set of K workers."""
    n = min(torch.cuda.device_count(), 4)
    if n < 3:
        pytest.skip("needs >= 3 GPUs")
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", str(n), "--timeout", "200", os.path.join(ROOT, "tests", "mp_litmus.py"), "ps", "40", "push", "resnet18", str(n - 1)]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=260)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _cli(n, d, extra, timeout=300):
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", str(n), "--timeout", str(timeout - 40), os.path.join(ROOT, "src", "distributed_nn.py"),
           "--network", "ResNet18", "--dataset", "Cifar10", "--batch-size", "32", "--lr", "0.05", "--momentum", "0.9",
           "--train-dir", d, "--log-interval", "1", "--enable-gpu", "true"] + extra
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=timeout)
    return r


@pytest.mark.parametrize("mode", ["kill", "timeout"])
def test_cli_straggler_modes_on_nvlink(tmp_path, mode):
    """--mode kill / timeout on the NVLink tier with an injected straggler: the job finishes, the PS never waits for the slow worker
    (kill: its abort word is raised by the PS kernel and it abandons the step; timeout: the PS proceeds with what arrived), loss stays finite."""
    n = min(torch.cuda.device_count(), 4)
    if n < 3:
        pytest.skip("needs >= 3 GPUs")
    d = str(tmp_path) + "/"
    # worker co-located on the PS rank: P = n workers, the last one sleeps 400 ms before every step
    extra = ["--max-steps", "8", "--eval-freq", "1000", "--mode", mode, "--inject-straggler", "%d:400" % (n - 1), "--cuda-graph", "false",
             "--colocate-worker", "true"]
    extra += ["--num-aggregate", str(n - 1)] if mode == "kill" else ["--num-aggregate", str(n), "--kill-threshold", "0.15"]
    r = _cli(n, d, extra)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    from ps_pytorch_b200.utils.obs import parse_worker_line
    recs = [x for x in (parse_worker_line(l) for l in (r.stdout + r.stderr).splitlines()) if x and x["rank"] == 0]
    assert len(recs) == 8 and all(x["loss"] == x["loss"] for x in recs)
    # the PS must not have been slowed to the straggler's pace (8 steps x 0.4 s = 3.2 s if it had waited for it every step)
    ps_lines = [l for l in (r.stdout + r.stderr).splitlines() if "Master node is entering step" in l]
    assert len(ps_lines) >= 8


def test_cli_resume_on_nvlink(tmp_path):
    """--resume on the NVLink tier: a second launch continues from the checkpoint step; device step counters are seeded with it."""
    n = min(torch.cuda.device_count(), 2)
    d = str(tmp_path) + "/"
    r = _cli(n, d, ["--max-steps", "6", "--eval-freq", "3"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    from ps_pytorch_b200.utils.checkpoint import latest_step
    assert latest_step(d) == 6
    r = _cli(n, d, ["--max-steps", "10", "--eval-freq", "5", "--resume", "true"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "resumed from step 6" in (r.stdout + r.stderr)
    assert latest_step(d) == 10
    steps = [int(l.rsplit(":", 1)[1]) for l in (r.stdout + r.stderr).splitlines() if "Master node is entering step" in l]
    assert steps and steps[0] == 7 and steps[-1] == 10, steps


def test_nvl_overlap_with_graph_matches_plain(tmp_path):
    """--overlap + CUDA graph: bucket pushes recorded inside the graph must give the same weights as the plain path.
    Two steps only: the weight-gradient kernels sum their split-K partials with fp32 atomics, so two runs of the SAME configuration drift apart
    chaotically (batch-32 BatchNorm, lr 0.1): measured cos 0.97 between identical plain runs after 4 steps (scripts/overlap_noise.py,
    profiles/run27_multi_gpu.md) — a lost or stale bucket is an O(1) effect in the first two steps, the noise is not."""
    n = min(torch.cuda.device_count(), 2)
    base = ["--steps", "2", "--network", "ResNet18", "--bs", "32", "--colocate", "1", "--backend", "sm100"]
    a = _run(n, str(tmp_path / "a.pt"), base + ["--graph", "1", "--overlap", "1"])
    b = _run(n, str(tmp_path / "b.pt"), base + ["--graph", "0", "--overlap", "0"])
    init = _init("ResNet18")
    da = torch.cat([(x - i).flatten() for x, i in zip(a["w"], init)])
    db = torch.cat([(x - i).flatten() for x, i in zip(b["w"], init)])
    cos = torch.nn.functional.cosine_similarity(da, db, dim=0).item()
    assert cos > 0.95 and abs(da.norm().item() / db.norm().item() - 1.0) < 0.1, (cos, da.norm().item(), db.norm().item())


def test_nvl_async_pull_matches_reference(tmp_path):
    """--comm-type Async: workers pull the weights from the PS replica instead of receiving a broadcast."""
    n = min(torch.cuda.device_count(), 2)
    res = _run(n, str(tmp_path / "w.pt"), ["--steps", "3", "--network", "LeNet", "--bs", "64", "--colocate", "1", "--backend", "library", "--comm-type", "Async"])
    _check(res, _reference(n, 3, 64, "LeNet"), _init("LeNet"))


DDP_SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from ps_pytorch_b200.models import build_model
from ps_pytorch_b200.parallel.ddp import DistributedDataParallel
r, w = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("gloo", rank=r, world_size=w)
torch.manual_seed(100 + r)
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
ddp = DistributedDataParallel(build_model("ResNet18", 10).to(torch.bfloat16), dev)
ok = True
for it in range(3):
    g = torch.Generator().manual_seed(5 + r + 10 * it)
    x = torch.randn(16, 3, 32, 32, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (16,), generator=g).to(dev)
    ddp.zero_grad()
    torch.nn.functional.cross_entropy(ddp(x).float(), y).backward()
    local = ddp.grads.flat.float().clone()
    ddp.sync_gradients()
    torch.cuda.synchronize()
    # reference: gloo all-reduce of the local gradients
    ref = local.cpu(); dist.all_reduce(ref); ref /= w
    got = ddp.grads.flat.float().cpu()
    err = (got - ref).abs().max().item() / max(1e-6, ref.abs().max().item())
    ok = ok and err < 2e-2
    with torch.no_grad():
        ddp.weights.flat.add_(ddp.grads.flat, alpha=-0.01)
ddp._nvl.check_error()
t = torch.tensor([0 if ok else 1]); dist.all_reduce(t)
if r == 0: print("DDP_OK" if int(t[0]) == 0 else "DDP_BAD")
dist.barrier()
'''


def test_ddp_allreduce_kernel(tmp_path):
    n = min(torch.cuda.device_count(), 4)
    script = tmp_path / "ddp_gpu.py"
    script.write_text(DDP_SCRIPT % ROOT)
    r = subprocess.run([sys.executable, "-m", "ps_pytorch_b200.launch", "-n", str(n), "--timeout", "200", str(script)],
                       cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=260)
    assert r.returncode == 0 and "DDP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("mode", [["--topology", "sharded"], ["--parallelism", "ddp"]])
def test_cli_sharded_and_ddp_on_gpus(tmp_path, mode):
    """The two non-reference parallel modes through the reference-shaped CLI on real GPUs: loss falls, rank 0 writes a loadable checkpoint."""
    n = min(torch.cuda.device_count(), 2)
    d = str(tmp_path) + "/"
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", str(n), "--timeout", "240", os.path.join(ROOT, "src", "distributed_nn.py"),
           "--network", "ResNet18", "--dataset", "Cifar10", "--batch-size", "64", "--max-steps", "12", "--lr", "0.05", "--momentum", "0.9",
           "--eval-freq", "12", "--train-dir", d, "--log-interval", "1", "--enable-gpu", "true"] + mode
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    from ps_pytorch_b200.utils.obs import parse_worker_line
    recs = [x for x in (parse_worker_line(l) for l in (r.stdout + r.stderr).splitlines()) if x and x["rank"] == 0]
    assert len(recs) == 12 and recs[-1]["loss"] < recs[0]["loss"], [x["loss"] for x in recs]
    from ps_pytorch_b200.models import build_model
    from ps_pytorch_b200.utils.checkpoint import load_model_into, model_path
    m = build_model("ResNet18", 10)
    load_model_into(m, model_path(d, 12))
    assert all(torch.isfinite(p).all() for p in m.parameters())


def test_nvls_ld_reduce_sharded_matches_reference(tmp_path):
    """Sharded topology with the multimem.ld_reduce variant of the reduce kernel must produce the reference update."""
    n = min(torch.cuda.device_count(), 2)
    res = _run(n, str(tmp_path / "w.pt"), ["--steps", "3", "--network", "LeNet", "--bs", "64", "--topology", "sharded", "--grad-dtype", "bf16",
                                           "--colocate", "1", "--backend", "library"], env_extra={"PSB200_NVLS_REDUCE": "1"})
    ref = _reference(n, 3, 64, "LeNet")
    _check(res, ref, _init("LeNet"), 0.97)
