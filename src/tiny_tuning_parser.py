"""Average the last-step loss over workers for one learning-rate candidate (reference ``src/tiny_tuning_parser.py``).
Parses the *current* worker log line (the reference's regex expects an older format and matches nothing)."""
import argparse

import _path  # noqa: F401
from ps_pytorch_b200.utils.obs import parse_worker_line

parser = argparse.ArgumentParser(description='Distributed Tunning')
parser.add_argument('--tuning-dir', type=str, default='./tune/', metavar='N', help='file with the grepped last-step worker lines')
parser.add_argument('--tuning-lr', type=float, default=0.125, metavar='N', help='candidate learning rate used during tunning')
parser.add_argument('--num-workers', type=int, default=16, help='number of workers in the cluster')
args = parser.parse_args()

loss_stat = []
with open(args.tuning_dir, 'r') as file:
    for line in file.readlines():
        d = parse_worker_line(line.rstrip('\n'))
        if d:
            loss_stat.append(d["loss"])
if len(loss_stat) != args.num_workers:
    print("Illegal Number of Workers! (found {}, expected {})".format(len(loss_stat), args.num_workers))
if loss_stat:
    print("Avged loss for lr candidate: {}=========>{}".format(args.tuning_lr, sum(loss_stat) / float(len(loss_stat))))
