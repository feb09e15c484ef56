"""ps_pytorch_b200 — a Blackwell-native parameter-server training engine.

Capabilities follow hwang595/ps_pytorch: rank 0 holds the global
model (``SyncReplicasMaster_NN``), ranks 1..P are ``DistributedWorker`` replicas,
a ``DistributedEvaluator`` polls ``<train_dir>/model_step_<N>`` checkpoints.
The mechanism is new: flat parameter/gradient arenas in peer-mapped HBM, a
push kernel (cast/quantise + peer store + release flag), a PS-side
reduce+optimizer+broadcast kernel, and sm_100a compute kernels behind
``ps_pytorch_b200.nn`` layers.  A gloo transport is the no-GPU plumbing tier.
"""
__version__ = "0.1.0"

from .config import add_fit_args, add_eval_args, add_single_machine_args, str2bool  # noqa: F401
from .models import build_model  # noqa: F401
