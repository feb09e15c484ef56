"""Tooling around the engine: log analysis, LR-sweep parser, cluster operations on a localhost inventory, and the reference-named
module facades under src/ (what a script written against the reference imports)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")


def _write_log(path, nworkers, step_time, steps=12):
    from ps_pytorch_b200.utils.obs import format_worker_line
    with open(path, "w") as f:
        f.write("I am worker: 1 in all %d workers, next step: 1\n" % nworkers)
        for s in range(1, steps + 1):
            for r in range(1, nworkers + 1):
                f.write(format_worker_line(r, s, 0, s * 8, 1000, 2.0 / s, step_time * (1.0 + 0.1 * (r - 1)), 0.01, 0.1, 0.2, 0.05) + "\n")


def test_speedup_analysis_from_logs(tmp_path):
    l1, l4 = str(tmp_path / "w1.log"), str(tmp_path / "w4.log")
    _write_log(l1, 1, 0.40)
    _write_log(l4, 4, 0.125)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "analysis", "speedup.py"), "--steps", "10", "1:" + l1, "4:" + l4],
                       capture_output=True, text=True, env=ENV, timeout=120)
    assert r.returncode == 0, r.stderr[-1500:]
    rows = [l.split() for l in r.stdout.splitlines() if l.strip() and l.split()[0].isdigit()]
    by_n = {int(x[0]): x for x in rows}
    assert set(by_n) == {1, 4}
    # slowest of 4 workers takes 0.125 * 1.3 = 0.1625 s/step -> speed-up 0.40 / 0.1625 = 2.46; fastest -> 3.2 ("ideal")
    assert abs(float(by_n[4][3]) - 0.40 / 0.1625) < 0.05 and abs(float(by_n[4][4]) - 3.2) < 0.05
    assert abs(float(by_n[1][3]) - 1.0) < 1e-6


def test_tiny_tuning_parser(tmp_path):
    from ps_pytorch_b200.utils.obs import format_worker_line
    f = tmp_path / "lr_0.125_processing"
    f.write_text("\n".join(format_worker_line(r, 100, 0, 800, 1000, 0.5 + 0.1 * r, 0.1, 0, 0, 0, 0) for r in (1, 2, 3)) + "\n")
    r = subprocess.run([sys.executable, "tiny_tuning_parser.py", "--tuning-dir", str(f), "--tuning-lr", "0.125", "--num-workers", "3"],
                       cwd=os.path.join(ROOT, "src"), capture_output=True, text=True, env=ENV, timeout=120)
    assert r.returncode == 0, r.stderr[-1500:]
    assert "Illegal" not in r.stdout and "0.125=========>0.7" in r.stdout


def test_cluster_tool_on_localhost(tmp_path):
    inv = tmp_path / "inventory"
    inv.write_text("localhost head\n127.0.0.1 second   # comment\n")
    tool = os.path.join(ROOT, "tools", "cluster.py")
    r = subprocess.run([sys.executable, tool, "get_hosts", "--inventory", str(inv), "--out-dir", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "hosts_alias").read_text().split() == ["head", "second"]
    assert (tmp_path / "hosts_address").read_text().split() == ["localhost", "127.0.0.1"]
    r = subprocess.run([sys.executable, tool, "run_command", "--inventory", str(inv), "--", "echo hello-$((40+2))"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.count("hello-42") == 2 and "=== head" in r.stdout
    r = subprocess.run([sys.executable, tool, "run_command", "--inventory", str(inv), "--", "exit 3"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3


def test_reference_named_modules_resolve():
    """`cd src; python -c 'from model_ops.resnet import ResNet18'` etc. — every module path a reference-era script imports."""
    code = """
import torch
from model_ops.lenet import LeNet, LeNetSplit
from model_ops.resnet import ResNet18, ResNet34, ResNet50, ResNet101, ResNet152
from model_ops.resnet_split import ResNetSplit18
from model_ops.vgg import vgg11_bn, vgg16, vgg19_bn
from optim.sgd import SGD
from optim.adam import Adam
from nn_ops import NN_Trainer, accuracy
from util import build_model, prepare_data
from sync_replicas_master_nn import SyncReplicasMaster_NN, GradientAccumulator
from distributed_worker import DistributedWorker, ModelBuffer
from distributed_evaluator import DistributedEvaluator
from data_loader_ops.my_data_loader import DataLoader
from compression import g_compress, g_decompress, w_compress, w_decompress
from data_parallel_dist import DistributedDataParallel
from distributed_functions import backward
m = build_model('ResNet18', 10)
assert sum(p.numel() for p in m.parameters()) == 11173962
out = LeNet()(torch.zeros(2, 1, 28, 28))
assert out.shape == (2, 10)
o = torch.randn(4, 10); t = o.argmax(1)
assert accuracy(o, t, topk=(1, 5))[0].item() == 100.0
print('facades ok')
"""
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.join(ROOT, "src"), capture_output=True, text=True, env=ENV, timeout=300)
    assert r.returncode == 0 and "facades ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_reference_launch_scripts(tmp_path):
    """src/run_pytorch.sh -> src/evaluate_pytorch.sh, the reference's two launch lines, on a small CPU configuration."""
    d = str(tmp_path) + "/"
    env = dict(ENV, NPROC="3", NETWORK="LeNet", DATASET="MNIST", BATCH="32", MAX_STEPS="6", EVAL_FREQ="3", TRAIN_DIR=d, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(["bash", os.path.join(ROOT, "src", "run_pytorch.sh")], capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.isfile(d + "model_step_3") and os.path.isfile(d + "model_step_6")
    env["EVAL_BATCH"] = "256"
    r = subprocess.run(["bash", os.path.join(ROOT, "src", "evaluate_pytorch.sh"), "--max-evals", "2", "--poll-seconds", "0.2"],
                       capture_output=True, text=True, env=env, timeout=200)
    assert r.returncode == 0 and r.stdout.count("Test set: Average loss") == 2, r.stdout[-1500:] + r.stderr[-1500:]


def test_lr_sweep_script(tmp_path):
    """src/tune.sh: run each learning-rate candidate for a few steps, grep the last step's worker lines, average the loss per candidate."""
    d = str(tmp_path) + "/"
    env = dict(ENV, NPROC="3", NETWORK="LeNet", DATASET="MNIST", BATCH="16", MAX_TUNING_STEP="4", TUNE_DIR=d, LRS="0.01 0.1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(["bash", os.path.join(ROOT, "src", "tune.sh")], capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("Avged loss for lr candidate")]
    assert len(lines) == 2 and "0.01=========>" in lines[0] and "0.1=========>" in lines[1], r.stdout[-1500:]
    assert "Illegal Number of Workers" not in r.stdout


def test_scaling_table_from_bench_json(tmp_path):
    """analysis/scaling.py: speed-up / efficiency against the smallest run, NCCL ratio where the bench measured the baseline in the same job."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = []
    for n, v, nb in ((1, 100.0, 50.0), (4, 380.0, None)):
        p = tmp_path / ("n%d.json" % n)
        p.write_text(json.dumps({"n_gpus": n, "value": v, "ms_per_step": 10.0, "scaling": "weak", "exposed_comm_ms": 0.1,
                                 "nccl_baseline": {"value": nb} if nb else None}))
        files.append(str(p))
    out = subprocess.run([sys.executable, os.path.join(root, "analysis", "scaling.py")] + files[::-1], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("| 1 ") or l.startswith("| 4 ")]
    assert "| 1.00 | 1.000 |" in lines[0] and "| 2.00 |" in lines[0]
    assert "| 3.80 | 0.950 |" in lines[1] and lines[1].rstrip().endswith("| - | - |")
