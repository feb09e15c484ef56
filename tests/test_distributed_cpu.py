"""Multi-process plumbing on the gloo tier (the designated fake backend): PS(P workers) must equal
single-process SGD on the concatenated batch; K-of-P; lossless codec on the wire; co-located worker."""
import os
import subprocess
import sys

import pytest
import torch

from ps_pytorch_b200.models import build_model
from ps_pytorch_b200.optim import SGD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nprocs, out, extra, timeout=240):
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", str(nprocs), "--timeout", str(timeout - 20),
           os.path.join(ROOT, "tests", "mp_job.py"), "--out", out] + extra
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return torch.load(out)


def _reference(nworkers, steps, bs, network="LeNet"):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mp_job import batch_for
    torch.manual_seed(7)
    m = build_model(network, 10)
    opt = SGD(m.parameters(), lr=0.05, momentum=0.9)
    shape = (1, 28, 28) if network.startswith("LeNet") else (3, 32, 32)
    for s in range(1, steps + 1):
        xs, ys = zip(*[batch_for(k, s, bs, shape, 10) for k in range(nworkers)])
        m.zero_grad()
        # mean of per-worker mean losses == PS average of worker gradients
        loss = sum(torch.nn.functional.cross_entropy(m(x), y) for x, y in zip(xs, ys)) / nworkers
        loss.backward()
        opt.step(grads=[p.grad for p in m.parameters()])
    return [p.detach().clone() for p in m.parameters()]


@pytest.mark.parametrize("compress,colocate", [(0, 0), (1, 0), (0, 1)])
def test_ps_equals_single_process_sgd(tmp_path, compress, colocate):
    res = _run(3, str(tmp_path / "w.pt"), ["--steps", "3", "--compress", str(compress), "--colocate", str(colocate)])
    nw = 3 if colocate else 2
    assert res["workers"] == nw
    ref = _reference(nw, 3, 8)
    for a, b in zip(res["w"], ref):
        assert torch.allclose(a, b, atol=2e-5, rtol=1e-4)


def test_k_of_p_runs_and_uses_k(tmp_path):
    res = _run(4, str(tmp_path / "w.pt"), ["--steps", "3", "--k", "2"])
    assert res["workers"] == 3 and all(torch.isfinite(w).all() for w in res["w"])
    ref = _reference(3, 3, 8)
    # with K=2 of 3 the update differs from the all-worker mean
    assert any(not torch.allclose(a, b, atol=1e-6) for a, b in zip(res["w"], ref))


def test_overlap_buckets_same_result(tmp_path):
    a = _run(3, str(tmp_path / "a.pt"), ["--steps", "2", "--overlap", "1"])
    ref = _reference(2, 2, 8)
    for x, y in zip(a["w"], ref):
        assert torch.allclose(x, y, atol=2e-5, rtol=1e-4)


def test_cli_end_to_end_with_evaluator(tmp_path):
    d = str(tmp_path) + "/"
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", "3", "--timeout", "200", os.path.join(ROOT, "src", "distributed_nn.py"),
           "--network", "LeNet", "--dataset", "MNIST", "--batch-size", "32", "--max-steps", "10", "--lr", "0.05", "--momentum", "0.9",
           "--num-aggregate", "2", "--eval-freq", "5", "--train-dir", d, "--compress-grad", "compress", "--log-interval", "5",
           "--mode", "normal", "--comm-type", "Bcast", "--enable-gpu="]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.isfile(d + "model_step_5") and os.path.isfile(d + "model_step_10") and os.path.isfile(d + "model_step_10.opt")
    assert "Worker: 1, Step: 10" in r.stdout + r.stderr
    ev = subprocess.run([sys.executable, os.path.join(ROOT, "src", "distributed_evaluator.py"), "--model-dir", d, "--eval-freq", "5",
                         "--max-evals", "2", "--poll-seconds", "0.2", "--eval-batch-size", "512"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=200)
    assert ev.returncode == 0 and ev.stdout.count("Test set: Average loss") == 2, ev.stdout + ev.stderr


def test_kill_mode_with_injected_straggler(tmp_path):
    d = str(tmp_path) + "/"
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", "4", "--timeout", "200", os.path.join(ROOT, "src", "distributed_nn.py"),
           "--network", "LeNetSplit", "--dataset", "MNIST", "--batch-size", "16", "--max-steps", "6", "--num-aggregate", "2",
           "--eval-freq", "100", "--train-dir", d, "--compress-grad", "None", "--mode", "kill", "--inject-straggler", "3:400"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "Master node is entering step: 6" in r.stdout + r.stderr


def test_single_machine_cli(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "src", "single_machine.py"), "--network", "LeNet", "--dataset", "MNIST",
                        "--batch-size", "64", "--epochs", "1", "--max-steps", "25", "--lr", "0.05", "--momentum", "0.9", "--no-cuda"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    losses = [float(l.split("Loss: ")[1].split()[0]) for l in r.stdout.splitlines() if l.startswith("Train Epoch")]
    assert len(losses) >= 2 and losses[-1] < losses[0]
    assert "Test set: Average loss" in r.stdout


def test_resume_from_checkpoint_sidecar(tmp_path):
    """Train 6 steps (checkpoints + .opt sidecars at 3 and 6), then --resume to step 9: the PS restarts at step 7 with the
    saved momentum, and the result equals an uninterrupted 9-step run (the reference has no resume at all)."""
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")
    base = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", "2", "--timeout", "200", os.path.join(ROOT, "src", "distributed_nn.py"),
            "--network", "LeNet", "--dataset", "MNIST", "--batch-size", "32", "--lr", "0.05", "--momentum", "0.9", "--num-aggregate", "1",
            "--eval-freq", "3", "--compress-grad", "None", "--seed", "3"]
    da, db = str(tmp_path / "a") + "/", str(tmp_path / "b") + "/"
    os.makedirs(da); os.makedirs(db)
    r = subprocess.run(base + ["--max-steps", "9", "--train-dir", da], cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run(base + ["--max-steps", "6", "--train-dir", db], cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run(base + ["--max-steps", "9", "--train-dir", db, "--resume", "true"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "resumed from step 6" in r.stdout + r.stderr
    assert "Master node is entering step: 7" in r.stdout + r.stderr and "entering step: 6\n" not in r.stdout + r.stderr
    assert os.path.isfile(db + "model_step_9")


def test_fail_stop_recovery_with_restarts(tmp_path):
    """A worker dies at step 5 (fault injection); the supervising launcher tears the job down, restarts it with --resume, the PS reloads
    model_step_3 + its optimizer sidecar and training continues to step 9 (the reference has neither failure detection nor resume)."""
    d = str(tmp_path) + "/"
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", "3", "--timeout", "200", "--max-restarts", "2", os.path.join(ROOT, "src", "distributed_nn.py"),
           "--network", "LeNet", "--dataset", "MNIST", "--batch-size", "16", "--max-steps", "9", "--eval-freq", "3", "--train-dir", d,
           "--compress-grad", "None", "--inject-failure", "2:5", "--log-interval", "1", "--enable-gpu="]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=280)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "injected failure at step 5" in out and "restart 1/2 with --resume" in out
    assert "resumed from step 3" in out
    assert os.path.isfile(d + "model_step_9")
    assert "Worker: 1, Step: 9" in out
