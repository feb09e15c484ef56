"""Entry point of the distributed job — same script name and flags as the reference
(``/root/reference/src/distributed_nn.py``).  Launch with
``python -m ps_pytorch_b200.launch -n N src/distributed_nn.py --flags`` (or torchrun); rank 0 becomes the
parameter server, the other ranks workers."""
import argparse

import _path  # noqa: F401
from ps_pytorch_b200.config import add_fit_args
from ps_pytorch_b200.parallel.runtime import run_distributed

from distributed_worker import *  # noqa: F401,F403  (reference re-exports)
from sync_replicas_master_nn import *  # noqa: F401,F403

if __name__ == "__main__":
    args = add_fit_args(argparse.ArgumentParser(description='B200 parameter-server training'))
    run_distributed(args)
