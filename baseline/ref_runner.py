"""Per-process entry used by ``baseline/ref_arm.py``: runs ONE rank of the unmodified reference.

    python ref_runner.py <reference src dir> <script.py> [script args...]

What it changes is the environment around the reference, never the reference's code:
  * ``mpi4py`` / ``blosc`` resolve to the stand-ins in ``baseline/refshim`` (neither is installed; no network);
  * torchvision's md5 integrity check is made to accept the synthetic dataset files written by ``ref_arm.py``
    (the reference asks torchvision to download CIFAR-10/MNIST; there is no network);
  * a logging handler timestamps the reference's own "Master node is entering step" / worker step log lines.
"""
import json
import logging
import os
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


class _StepClock(logging.Handler):
    def __init__(self, path):
        logging.Handler.__init__(self)
        self.f = open(path, "a", buffering=1)

    def emit(self, record):
        try:
            msg = record.getMessage()
        except Exception:
            return
        if msg.startswith("Master node is entering step:"):
            self.f.write(json.dumps({"t": time.time(), "ps_step": int(msg.rsplit(":", 1)[1])}) + "\n")
        elif msg.startswith("Worker:") and "Loss:" in msg:
            try:
                step = int(msg.split("Step:")[1].split(",")[0])
                loss = float(msg.split("Loss:")[1].split(",")[0])
                self.f.write(json.dumps({"t": time.time(), "worker_step": step, "loss": loss, "rank": int(os.environ.get("REFSHIM_RANK", 0)),
                                         "line": msg}) + "\n")
            except Exception:
                pass


def main():
    ref_src, script = sys.argv[1], sys.argv[2]
    sys.argv = [os.path.join(ref_src, script)] + sys.argv[3:]
    sys.path.insert(0, ref_src)
    sys.path.insert(0, os.path.join(HERE, "refshim"))
    import torchvision.datasets.utils as tvu
    tvu.check_integrity = lambda fpath, md5=None: os.path.isfile(fpath)
    for modname in ("cifar", "mnist", "svhn"):
        try:
            mod = __import__("torchvision.datasets." + modname, fromlist=["x"])
            if hasattr(mod, "check_integrity"):
                mod.check_integrity = tvu.check_integrity
        except Exception:
            pass
    log = os.environ.get("REFSHIM_TIMELOG")
    if log:
        logging.getLogger().addHandler(_StepClock(log))
    threads = os.environ.get("REFSHIM_TORCH_THREADS")
    if threads:
        import torch
        torch.set_num_threads(int(threads))
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
