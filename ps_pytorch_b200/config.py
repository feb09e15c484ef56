"""One flag schema shared by every entry point.

Flag names and defaults are those of the reference parsers
(``/root/reference/src/distributed_nn.py:24-68``, ``single_machine.py:24-51``,
``distributed_evaluator.py:39-56``); the superset flags (``--transport``,
``--grad-dtype`` …) are new.  ``--enable-gpu`` is parsed as a real boolean here
(the reference uses ``type=bool`` so any non-empty string is True; ``--enable-gpu=``
still means False, which keeps ``run_pytorch.sh``-style launch lines working).
"""
from __future__ import annotations

import argparse

NETWORKS = ["LeNet", "LeNetSplit", "ResNet18", "ResNet34", "ResNet50", "ResNet101", "ResNet152",
            "ResNetSplit18", "ResNet50ImageNet", "ResNet18ImageNet",
            "VGG11", "VGG11_bn", "VGG13", "VGG13_bn", "VGG16", "VGG16_bn", "VGG19", "VGG19_bn"]
DATASETS = ["MNIST", "Cifar10", "Cifar100", "SVHN", "ImageNet"]


def str2bool(v) -> bool:
    if isinstance(v, bool):
        return v
    s = str(v).strip().lower()
    if s in ("", "0", "false", "f", "no", "n", "off", "none"):
        return False
    if s in ("1", "true", "t", "yes", "y", "on"):
        return True
    raise argparse.ArgumentTypeError("expected a boolean, got %r" % (v,))


def _common_train(parser: argparse.ArgumentParser, test_bs: int) -> None:
    parser.add_argument('--batch-size', type=int, default=128, metavar='N', help='per-worker training batch size')
    parser.add_argument('--test-batch-size', type=int, default=test_bs, metavar='N', help='batch size for testing')
    parser.add_argument('--epochs', type=int, default=100, metavar='N', help='number of epochs to train')
    parser.add_argument('--lr', type=float, default=0.01, metavar='LR', help='learning rate')
    parser.add_argument('--momentum', type=float, default=0.5, metavar='M', help='SGD momentum')
    parser.add_argument('--no-cuda', action='store_true', default=False, help='disables CUDA training')
    parser.add_argument('--seed', type=int, default=1, metavar='S', help='random seed (honoured here; ignored by the reference)')
    parser.add_argument('--log-interval', type=int, default=10, metavar='N', help='log every N steps')
    parser.add_argument('--network', type=str, default='LeNet', metavar='N', help='one of: ' + ', '.join(NETWORKS))
    parser.add_argument('--dataset', type=str, default='MNIST', metavar='N', help='one of: ' + ', '.join(DATASETS))
    # superset
    parser.add_argument('--synthetic', type=str2bool, default=None, nargs='?', const=True,
                        help='use synthetic data of the dataset shape (default: auto when the dataset is not on disk)')
    parser.add_argument('--data-root', type=str, default='.', help='where <dataset>_data directories live')
    parser.add_argument('--optimizer', type=str, default='sgd', choices=['sgd', 'adam'], help='PS-side optimizer')
    parser.add_argument('--weight-decay', type=float, default=0.0)
    parser.add_argument('--nesterov', type=str2bool, default=False, nargs='?', const=True)
    parser.add_argument('--dampening', type=float, default=0.0)
    parser.add_argument('--amsgrad', type=str2bool, default=False, nargs='?', const=True)
    parser.add_argument('--dtype', type=str, default='auto', choices=['auto', 'fp32', 'bf16'],
                        help='worker compute dtype (auto: bf16 on GPU, fp32 on CPU)')
    parser.add_argument('--backend', type=str, default='auto', choices=['auto', 'sm100', 'library'],
                        help='layer kernels: hand-written sm_100a, or the cuDNN/cuBLAS library path (baseline)')


def add_fit_args(parser: argparse.ArgumentParser, argv=None):
    """Flags of ``distributed_nn.py`` (reference ``src/distributed_nn.py:24-68``) + superset."""
    _common_train(parser, test_bs=500)
    parser.add_argument('--max-steps', type=int, default=10000, metavar='N', help='the maximum number of iterations')
    parser.add_argument('--mode', type=str, default='normal', metavar='N',
                        help="normal | kill (signal-kill stragglers) | timeout (kill after --kill-threshold seconds)")
    parser.add_argument('--kill-threshold', type=float, default=7.0, metavar='KT', help='straggler timeout in seconds')
    parser.add_argument('--topology', type=str, default='ps', choices=['ps', 'sharded'],
                        help='ps: rank 0 is the parameter server (reference) | sharded: every rank serves 1/W of the parameters (NVLink tier only)')
    parser.add_argument('--parallelism', type=str, default='ps', choices=['ps', 'ddp'],
                        help='ps: parameter server (the reference\'s product path) | ddp: all-reduce data parallelism (its vendored DistributedDataParallel)')
    parser.add_argument('--comm-type', type=str, default='Bcast', metavar='N', help='Bcast (multicast weight push) | Async (workers pull weights)')
    parser.add_argument('--num-aggregate', type=int, default=5, metavar='N',
                        help='K: how many gradients the PS aggregates per step (K-of-P backup workers); clamped to P')
    parser.add_argument('--eval-freq', type=int, default=50, metavar='N', help='checkpoint every N steps')
    parser.add_argument('--train-dir', type=str, default='output/models/', metavar='N', help='checkpoint directory (must end in /)')
    parser.add_argument('--compress-grad', type=str, default='compress', metavar='N',
                        help="compress | None — 'compress' = block-scaled fp8 on NVLink, lossless shuffle+LZ on gloo")
    parser.add_argument('--enable-gpu', type=str2bool, default=False, nargs='?', const=True, help='run workers (and the PS) on GPUs')
    # superset
    parser.add_argument('--transport', type=str, default='auto', choices=['auto', 'gloo', 'nvl', 'local'],
                        help='nvl: peer-memory kernels over NVLink; gloo: CPU plumbing tier; local: single process')
    parser.add_argument('--grad-dtype', type=str, default='auto', choices=['auto', 'fp32', 'bf16', 'fp8'],
                        help='wire format of pushed gradients (auto: fp8 if --compress-grad=compress else bf16 on GPU; fp32 on gloo)')
    parser.add_argument('--colocate-worker', type=str2bool, default=False, nargs='?', const=True,
                        help='also run a worker on the PS GPU (extension; reference leaves the PS device idle)')
    parser.add_argument('--overlap', type=str2bool, default=False, nargs='?', const=True,
                        help='push gradient buckets while backward is still running (the Split-model capability)')
    parser.add_argument('--num-buckets', type=int, default=4, help='bucket count for --overlap')
    parser.add_argument('--cuda-graph', type=str2bool, default=True, nargs='?', const=True, help='capture the worker step in a CUDA graph')
    parser.add_argument('--inject-straggler', type=str, default='', help='rank:ms[,rank:ms] artificial delay per step (fault injection)')
    parser.add_argument('--metrics-port', type=int, default=0, help='Prometheus endpoint: worker rank r serves http://0.0.0.0:<port+r>/metrics (0 = off)')
    parser.add_argument('--inject-failure', type=str, default='', help='rank:step — that worker exits with code 17 at that step on the first attempt (fault injection for --max-restarts)')
    parser.add_argument('--resume', type=str2bool, default=False, nargs='?', const=True, help='resume from the newest checkpoint + sidecar in --train-dir')
    parser.add_argument('--strict-reference', type=str2bool, default=False, nargs='?', const=True,
                        help='reproduce reference quirks (K+1 gate with divide-by-K)')
    parser.add_argument('--jsonl', type=str, default='', help='also write per-step metrics as JSON lines to this file')
    return parser.parse_args(argv)


def add_single_machine_args(parser: argparse.ArgumentParser, argv=None):
    """Flags of ``single_machine.py`` (reference ``src/single_machine.py:24-51``) + superset."""
    _common_train(parser, test_bs=1000)
    parser.add_argument('--max-steps', type=int, default=0, help='stop after N steps (0 = run all epochs)')
    parser.add_argument('--enable-gpu', type=str2bool, default=None, nargs='?', const=True, help='default: use the GPU when present and not --no-cuda')
    return parser.parse_args(argv)


def add_eval_args(parser: argparse.ArgumentParser, argv=None):
    """Flags of ``distributed_evaluator.py`` (reference ``src/distributed_evaluator.py:39-56``) + superset."""
    parser.add_argument('--eval-batch-size', type=int, default=10000, metavar='N')
    parser.add_argument('--eval-freq', type=int, default=50, metavar='N')
    parser.add_argument('--model-dir', type=str, default='output/models/', metavar='N')
    parser.add_argument('--dataset', type=str, default='MNIST', metavar='N')
    parser.add_argument('--network', type=str, default='LeNet', metavar='N')
    parser.add_argument('--synthetic', type=str2bool, default=None, nargs='?', const=True)
    parser.add_argument('--data-root', type=str, default='.')
    parser.add_argument('--enable-gpu', type=str2bool, default=False, nargs='?', const=True)
    parser.add_argument('--poll-seconds', type=float, default=10.0, help='sleep when the next checkpoint is absent')
    parser.add_argument('--max-evals', type=int, default=0, help='exit after N evaluations (0 = poll forever like the reference)')
    parser.add_argument('--timeout-seconds', type=float, default=0.0, help='give up after this long without a new checkpoint (0 = never)')
    return parser.parse_args(argv)


def num_classes_for(dataset: str) -> int:
    return {"Cifar100": 100, "ImageNet": 1000}.get(dataset, 10)


def input_shape_for(dataset: str):
    return {"MNIST": (1, 28, 28), "ImageNet": (3, 224, 224)}.get(dataset, (3, 32, 32))
