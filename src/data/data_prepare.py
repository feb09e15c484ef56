"""Prepare datasets once, before launching ranks (reference ``src/data/data_prepare.py``).

With network access this downloads MNIST / CIFAR-10 / CIFAR-100 / SVHN through torchvision into ``./<name>_data``.  On the
air-gapped B200 pod (no egress) it instead materialises the synthetic datasets of the same shapes under
``./<name>_data/synthetic.pt`` so every rank can ``torch.load`` them instead of regenerating."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from ps_pytorch_b200.utils.data import _DIRS, synthetic_dataset  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default=".")
    ap.add_argument("--download", action="store_true", help="try torchvision downloads (needs network)")
    ap.add_argument("--datasets", default="MNIST,Cifar10,Cifar100,SVHN")
    a = ap.parse_args()
    for name in a.datasets.split(","):
        d = os.path.join(a.root, _DIRS[name])
        os.makedirs(d, exist_ok=True)
        done = False
        if a.download:
            try:
                from torchvision import datasets
                if name == "MNIST":
                    datasets.MNIST(d, train=True, download=True); datasets.MNIST(d, train=False, download=True)
                elif name == "Cifar10":
                    datasets.CIFAR10(d, train=True, download=True); datasets.CIFAR10(d, train=False, download=True)
                elif name == "Cifar100":
                    datasets.CIFAR100(d, train=True, download=True); datasets.CIFAR100(d, train=False, download=True)
                elif name == "SVHN":
                    datasets.SVHN(d, split='train', download=True); datasets.SVHN(d, split='test', download=True)
                done = True
            except Exception as e:  # noqa: BLE001
                print("download of %s failed (%s); writing the synthetic stand-in" % (name, e))
        if not done:
            tr, te = synthetic_dataset(name, 16384, seed=0), synthetic_dataset(name, 2048, seed=1)
            torch.save({"train": (tr.images, tr.labels), "test": (te.images, te.labels)}, os.path.join(d, "synthetic.pt"))
        print("prepared", name, "->", d)
