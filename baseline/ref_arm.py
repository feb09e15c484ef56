"""``bench.py --impl reference``: run the UNMODIFIED reference (hwang595/ps_pytorch, copied verbatim into the git-ignored
``baseline/_ref``) through its own entry point ``src/distributed_nn.py`` — 1 parameter server (CPU, as the reference
forces, sync_replicas_master_nn.py:131) + one worker per GPU — and time its steps.

The reference is launched the way ``src/run_pytorch.sh`` launches it (same flags), except that ``mpirun`` is replaced by
this launcher + the ``baseline/refshim`` mpi4py/blosc stand-ins, ``--enable-gpu`` is on, and CIFAR-10 is a synthetic
dataset written in torchvision's on-disk format (no network).  Step time = spacing of the PS's own "entering step" log
lines (wall clock: the reference's PS is host code, there is no device timeline to time)."""
from __future__ import annotations

import json
import os
import pickle
import shutil
import signal
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def ensure_reference():
    """Returns the reference's src dir, or raises RuntimeError(reason)."""
    src = os.path.join(REF, "src")
    if os.path.isfile(os.path.join(src, "distributed_nn.py")):
        return src
    if os.path.isdir("/root/reference/src"):
        os.makedirs(REF, exist_ok=True)
        for name in os.listdir("/root/reference"):
            s, d = os.path.join("/root/reference", name), os.path.join(REF, name)
            if name.startswith(".git"):
                continue
            if os.path.isdir(s):
                shutil.copytree(s, d, dirs_exist_ok=True)
            else:
                shutil.copy2(s, d)
        return src
    raise RuntimeError("baseline/_ref is empty and /root/reference is not mounted on this box (run baseline/install_ref.sh in the build container)")


def write_synthetic_cifar10(root, n_train=50000, n_test=10000, seed=0):
    """CIFAR-10 'python version' layout with class-separable synthetic images (a per-class mean image + noise)."""
    import numpy as np
    d = os.path.join(root, "cifar-10-batches-py")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(seed)
    means = rng.integers(40, 216, size=(10, 3072)).astype(np.float32)

    def make(n):
        labels = rng.integers(0, 10, size=n)
        data = means[labels] + rng.normal(0.0, 48.0, size=(n, 3072)).astype(np.float32)
        return np.clip(data, 0, 255).astype(np.uint8), [int(v) for v in labels]

    per = n_train // 5
    for i in range(1, 6):
        data, labels = make(per)
        with open(os.path.join(d, "data_batch_%d" % i), "wb") as f:
            pickle.dump({"data": data, "labels": labels}, f, protocol=2)
    data, labels = make(n_test)
    with open(os.path.join(d, "test_batch"), "wb") as f:
        pickle.dump({"data": data, "labels": labels}, f, protocol=2)
    with open(os.path.join(d, "batches.meta"), "wb") as f:
        pickle.dump({"label_names": ["c%d" % i for i in range(10)]}, f, protocol=2)


def write_synthetic_mnist(root, n_train=60000, n_test=10000, seed=0):
    """MNIST idx files under <root>/MNIST/raw (class-separable synthetic digits)."""
    import numpy as np
    raw = os.path.join(root, "MNIST", "raw")
    os.makedirs(raw, exist_ok=True)
    rng = np.random.default_rng(seed)
    means = rng.integers(0, 200, size=(10, 784)).astype(np.float32)

    def make(n, img_name, lab_name):
        labels = rng.integers(0, 10, size=n).astype(np.uint8)
        data = np.clip(means[labels] + rng.normal(0.0, 40.0, size=(n, 784)), 0, 255).astype(np.uint8)
        with open(os.path.join(raw, img_name), "wb") as f:
            f.write((2051).to_bytes(4, "big") + n.to_bytes(4, "big") + (28).to_bytes(4, "big") + (28).to_bytes(4, "big") + data.tobytes())
        with open(os.path.join(raw, lab_name), "wb") as f:
            f.write((2049).to_bytes(4, "big") + n.to_bytes(4, "big") + labels.tobytes())

    make(n_train, "train-images-idx3-ubyte", "train-labels-idx1-ubyte")
    make(n_test, "t10k-images-idx3-ubyte", "t10k-labels-idx1-ubyte")


def _read_log(path):
    rows = []
    try:
        with open(path) as f:
            for line in f:
                try:
                    rows.append(json.loads(line))
                except ValueError:
                    pass
    except OSError:
        pass
    return rows


def run(gpus, steps, warmup, network="ResNet18", dataset="Cifar10", batch_size=1024, use_gpu=True, timeout_s=3000.0, keep=False):
    """Runs the reference for warmup+steps+1 PS steps; returns the result dict (or {"unavailable": why})."""
    try:
        src = ensure_reference()
    except RuntimeError as e:
        return {"impl": "reference", "unavailable": str(e)}
    warmup = max(int(warmup), 1)
    work = tempfile.mkdtemp(prefix="refarm_")
    os.makedirs(os.path.join(work, "models"), exist_ok=True)
    if dataset == "Cifar10":
        write_synthetic_cifar10(os.path.join(work, "cifar10_data"))
    elif dataset == "MNIST":
        write_synthetic_mnist(os.path.join(work, "mnist_data"))
    else:
        return {"impl": "reference", "unavailable": "no synthetic writer for dataset " + dataset}
    nworkers = max(int(gpus), 1)
    total = warmup + steps + 1
    timelog = os.path.join(work, "steps.jsonl")
    args = ["--lr=0.1", "--momentum=0.9", "--max-steps=%d" % total, "--epochs=100", "--network=" + network, "--dataset=" + dataset,
            "--batch-size=%d" % batch_size, "--comm-type=Bcast", "--num-aggregate=%d" % nworkers, "--mode=normal",
            "--eval-freq=%d" % (total + 1000), "--compress-grad=compress", "--enable-gpu=" + ("True" if use_gpu else ""),
            "--train-dir=" + os.path.join(work, "models") + "/"]
    procs = []
    visible = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v] or [str(i) for i in range(max(gpus, 1))]
    ncpu = os.cpu_count() or 8
    for r in range(nworkers + 1):
        env = dict(os.environ)
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "GROUP_RANK", "LOCAL_WORLD_SIZE", "TORCHELASTIC_RUN_ID"):
            env.pop(k, None)
        env.update({"REFSHIM_RANK": str(r), "REFSHIM_SIZE": str(nworkers + 1), "REFSHIM_DIR": os.path.join(work, "sock"),
                    "REFSHIM_TIMELOG": timelog, "PYTHONUNBUFFERED": "1",
                    "REFSHIM_TORCH_THREADS": str(max(1, min(16, ncpu // (nworkers + 1))))})
        # torchrun exports OMP_NUM_THREADS=1 to its children; the reference's host-side PS must not be throttled by the way THIS arm was launched
        env["OMP_NUM_THREADS"] = env["REFSHIM_TORCH_THREADS"]
        env["CUDA_VISIBLE_DEVICES"] = "" if (r == 0 or not use_gpu) else visible[(r - 1) % len(visible)]
        out = open(os.path.join(work, "rank%d.log" % r), "w")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "ref_runner.py"), src, "distributed_nn.py"] + args,
                                      cwd=work, env=env, stdout=out, stderr=subprocess.STDOUT, start_new_session=True))
    t_start, err = time.time(), None
    t_first = None
    try:
        while True:
            rows = _read_log(timelog)
            ps = {r["ps_step"]: r["t"] for r in rows if "ps_step" in r}
            if total in ps:
                break
            dead = [i for i, p in enumerate(procs) if p.poll() is not None]
            if dead:
                err = "reference rank %d exited with code %s before step %d; log tail: %s" % (
                    dead[0], procs[dead[0]].returncode, total, _tail(os.path.join(work, "rank%d.log" % dead[0])))
                break
            if time.time() - t_start > timeout_s:
                err = "reference did not reach step %d within %.0f s (reached %d)" % (total, timeout_s, max(ps) if ps else 0)
                break
            time.sleep(0.05)
    finally:
        for p in procs:                         # the reference never exits on its own (worker loop / PS wait at max_steps)
            try:
                os.killpg(p.pid, signal.SIGKILL)
            except OSError:
                pass
        for p in procs:
            try:
                p.wait(timeout=10)
            except Exception:
                pass
    rows = _read_log(timelog)
    ps = {r["ps_step"]: r["t"] for r in rows if "ps_step" in r}
    losses = sorted((r["worker_step"], r["loss"]) for r in rows if r.get("rank") == 1 and "loss" in r)
    sample = next((r["line"] for r in rows if r.get("rank") == 1 and r.get("worker_step") == warmup + 1), None)
    if not keep:
        shutil.rmtree(work, ignore_errors=True)
    if err is not None or (warmup + 1) not in ps or total not in ps:
        return {"impl": "reference", "unavailable": (err or "step log incomplete").replace("\n", " | ")[:600]}
    sec = ps[total] - ps[warmup + 1]
    gb = nworkers * batch_size
    value = gb * steps / sec
    c, h, w = (3, 32, 32) if dataset == "Cifar10" else (1, 28, 28)
    return {"metric": "ResNet-18/CIFAR-10 images/sec (whole box, device-timed, max over ranks)" if network == "ResNet18" else network + " images/sec",
            "value": value, "unit": "images/s", "n_gpus": gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3 / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32 (the reference has no reduced-precision path)" if use_gpu else "fp32 (CPU)",
            "data": "synthetic (class-separable CIFAR-10-format files read by the reference's own torchvision pipeline), random-init weights",
            "impl": "reference",
            "config": {"model": network, "global_batch": gb, "per_worker_batch": batch_size, "seq_len": None,
                       "parallelism": "ps: 1 CPU parameter server + %d %s workers%s (reference topology; mpirun replaced by baseline/refshim)" % (
                           nworkers, "GPU" if use_gpu else "CPU", (" on %d GPUs" % gpus) if use_gpu else ""),
                       "entry": "baseline/_ref/src/distributed_nn.py, flags of src/run_pytorch.sh with --enable-gpu=True",
                       "timing": "wall clock between the PS's 'entering step' log lines %d and %d" % (warmup + 1, total),
                       "sample_worker_log": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": batch_size * (c * h * w * 4 + 8),
                    "d2h_bytes_per_step": 4, "note": "the reference's loop is end to end by construction (DataLoader -> .to(device) -> loss.item() every step)"},
            "gpu_launches": 0, "first_loss": losses[0][1] if losses else None, "final_loss": losses[-1][1] if losses else None}


def _tail(path, n=6):
    try:
        with open(path) as f:
            return " | ".join(line.strip() for line in f.readlines()[-n:])
    except OSError:
        return ""


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--network", default="ResNet18")
    ap.add_argument("--dataset", default="Cifar10")
    ap.add_argument("--batch-size", type=int, default=1024)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.gpus, a.steps, a.warmup, a.network, a.dataset, a.batch_size, use_gpu=not a.cpu, keep=a.keep)))
