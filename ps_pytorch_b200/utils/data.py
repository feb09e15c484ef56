"""Data pipeline.

``prepare_data(args)`` mirrors the reference (``/root/reference/src/util.py:21-106``): MNIST (normalise
0.1307/0.3081), CIFAR-10/100 (reflect-pad 4 → random crop 32 → h-flip → normalise), SVHN.  Differences, all
forced by running on one 8×B200 box with no network:

* **synthetic mode** — class-separable random images of the dataset's shape, generated in memory (the pod has no
  datasets and cannot download); chosen automatically when the dataset directory is absent.
* the batch pipeline is an in-memory loader (``BatchLoader``) that assembles whole batches with vectorised
  ops (or the C++ prefetcher in ``ops/csrc/host_ext.cpp`` when built) into *pinned* host buffers, because a
  B200 worker consumes >100 k CIFAR images/s — three orders of magnitude past a per-sample Python transform.
* like the reference every worker sees the full training set and shuffles independently (seeded per rank).

``DataLoader`` keeps the reference's TF-style ``next_batch()`` API (``data_loader_ops/my_data_loader.py:318-319``).
"""
from __future__ import annotations

import collections
import os
from typing import Optional, Tuple

import torch

from ..config import input_shape_for, num_classes_for

_MEAN_STD = {
    "MNIST": ((0.1307,), (0.3081,)),
    "Cifar10": (tuple(x / 255.0 for x in (125.3, 123.0, 113.9)), tuple(x / 255.0 for x in (63.0, 62.1, 66.7))),
    "Cifar100": (tuple(x / 255.0 for x in (125.3, 123.0, 113.9)), tuple(x / 255.0 for x in (63.0, 62.1, 66.7))),
    "SVHN": ((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)),
    "ImageNet": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
}
_DIRS = {"MNIST": "mnist_data", "Cifar10": "cifar10_data", "Cifar100": "cifar100_data", "SVHN": "svhn_data",
         "ImageNet": "imagenet_data"}


class TensorDataset(torch.utils.data.Dataset):
    """uint8 images [N,C,H,W] + int64 labels, resident in host memory."""

    def __init__(self, images_u8: torch.Tensor, labels: torch.Tensor, name: str):
        self.images, self.labels, self.name = images_u8, labels.long(), name
        self.epochs_completed = 0

    def __len__(self):
        return self.images.size(0)

    def __getitem__(self, i):
        mean, std = _MEAN_STD[self.name]
        x = self.images[i].float().div_(255.0)
        m = torch.tensor(mean).view(-1, 1, 1)
        s = torch.tensor(std).view(-1, 1, 1)
        return (x - m) / s, self.labels[i]


def synthetic_dataset(name: str, n: int, seed: int = 0) -> TensorDataset:
    """Class-separable synthetic images: each class has a fixed random template; samples are template + noise."""
    g = torch.Generator().manual_seed(1234 + seed)
    c, h, w = input_shape_for(name)
    nc = num_classes_for(name)
    templates = torch.rand(nc, c, h, w, generator=torch.Generator().manual_seed(99))
    labels = torch.randint(0, nc, (n,), generator=g)
    chunk, outs = 4096, []
    for i in range(0, n, chunk):
        lab = labels[i:i + chunk]
        x = templates[lab] * 0.6 + torch.rand(lab.numel(), c, h, w, generator=g) * 0.4
        outs.append((x * 255.0).round().clamp_(0, 255).to(torch.uint8))
    return TensorDataset(torch.cat(outs), labels, name)


def _load_real(name: str, root: str, train: bool) -> Optional[TensorDataset]:
    """Load a torchvision dataset that is already on disk into memory (no download: the pod has no network)."""
    d = os.path.join(root, _DIRS[name])
    if not os.path.isdir(d):
        return None
    try:
        from torchvision import datasets
        import numpy as np
        if name == "MNIST":
            ds = datasets.MNIST(d, train=train, download=False)
            return TensorDataset(ds.data.unsqueeze(1).contiguous(), ds.targets, name)
        if name in ("Cifar10", "Cifar100"):
            ds = (datasets.CIFAR10 if name == "Cifar10" else datasets.CIFAR100)(root=d, train=train, download=False)
            return TensorDataset(torch.from_numpy(ds.data).permute(0, 3, 1, 2).contiguous(), torch.tensor(ds.targets), name)
        if name == "SVHN":
            ds = datasets.SVHN(d, split='train' if train else 'test', download=False)
            return TensorDataset(torch.from_numpy(np.ascontiguousarray(ds.data)), torch.from_numpy(ds.labels), name)
    except Exception:
        return None
    return None


class BatchLoader:
    """Whole-batch in-memory loader: shuffle → gather → (reflect-pad 4 + random crop + flip) → normalise,
    written into a pinned host buffer.  Iterating yields ``(images_f32_or_bf16, labels)`` host tensors."""

    def __init__(self, dataset: TensorDataset, batch_size: int, shuffle: bool, augment: bool, seed: int = 0,
                 drop_last: bool = False, pin: bool = False, out_dtype: torch.dtype = torch.float32,
                 channels_last: bool = False, prefetch: int = 2):
        self.dataset, self.batch_size, self.shuffle, self.augment = dataset, batch_size, shuffle, augment
        self.drop_last, self.out_dtype, self.channels_last = drop_last, out_dtype, channels_last
        self.gen = torch.Generator().manual_seed(seed)
        mean, std = _MEAN_STD[dataset.name]
        self.mean = torch.tensor(mean).view(1, -1, 1, 1)
        self.inv_std = 1.0 / torch.tensor(std).view(1, -1, 1, 1)
        self.pin = pin and torch.cuda.is_available()
        self.prefetch = prefetch
        # Pinned output buffers are a POOL with completion tracking, not a blind ring: a buffer goes back to the free list only after the
        # consumer has moved `hold` batches past it AND the CUDA event recorded after its host->device copy (mark_h2d_issued) has
        # completed.  The consumer keeps at most `hold` batches (the current one and the look-ahead the worker stages early).
        self.hold = 2
        self._nbuf = self.hold + max(prefetch, 0) + 2
        self._pool = None
        self._free = None
        self._outstanding = collections.deque()
        self._entry_of = {}
        try:
            from ..ops import host_ext
            self._native = host_ext.get() if out_dtype == torch.float32 else None
        except Exception:
            self._native = None

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _out_buffers(self, b):
        """A free pinned (images, labels) pair for a batch of ``b`` samples; blocks until the consumer has retired one."""
        if not self.pin:
            return None
        if self._pool is None or self._pool[0][0].size(0) < b:
            import queue
            c, h, w = self.dataset.images.shape[1:]
            n = max(b, self.batch_size)
            mk = lambda: (torch.empty(n, c, h, w, dtype=self.out_dtype).contiguous(
                memory_format=torch.channels_last if self.channels_last else torch.contiguous_format).pin_memory(),
                torch.empty(n, dtype=torch.int64).pin_memory())
            self._pool = [mk() for _ in range(self._nbuf)]
            self._free = queue.Queue()
            for e in self._pool:
                self._free.put(e)
        e = self._free.get()
        out = e if e[0].size(0) == b else (e[0][:b], e[1][:b])      # a short last batch is a view of a full-size buffer
        self._entry_of[id(out[0])] = e
        return out

    def _retire(self, keep: int) -> None:
        """Return buffers to the free list until at most ``keep`` yielded batches are outstanding."""
        while len(self._outstanding) > keep:
            x = self._outstanding.popleft()
            e = self._entry_of.pop(id(x), None)
            if e is None:
                continue
            ev = getattr(x, "_psb_evt", None)
            if ev is not None:
                ev.synchronize()                       # the H2D copy that read this buffer has finished
                x._psb_evt = None
            elif torch.cuda.is_available() and torch.cuda.is_initialized():
                torch.cuda.synchronize()               # consumer did not say when its copy was issued: be safe
            self._free.put(e)

    def _track(self, item):
        if self.pin and item is not None and id(item[0]) in self._entry_of:
            self._retire(self.hold - 1)
            self._outstanding.append(item[0])
        return item

    def _make_batch_native(self, idx: torch.Tensor):
        """One multi-threaded C++ pass: gather + reflect-pad crop + flip + normalise into the (pinned) output buffer."""
        b = idx.numel()
        c, h, w = self.dataset.images.shape[1:]
        out = self._out_buffers(b)
        if out is None:
            x = torch.empty(b, c, h, w, dtype=torch.float32)
            if self.channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            out = (x, torch.empty(b, dtype=torch.int64))
        dx = dy = fl = None
        if self.augment:
            dx = torch.randint(0, 9, (b,), generator=self.gen)
            dy = torch.randint(0, 9, (b,), generator=self.gen)
            fl = torch.rand(b, generator=self.gen) < 0.5
        self._native.assemble_batch(self.dataset.images, idx.contiguous(), out[0], self.mean.view(-1).contiguous(),
                                    self.inv_std.view(-1).contiguous(), dx, dy, fl)
        torch.index_select(self.dataset.labels, 0, idx, out=out[1])
        return out

    def make_batch(self, idx: torch.Tensor):
        if self._native is not None:
            return self._make_batch_native(idx)
        x = self.dataset.images[idx]
        y = self.dataset.labels[idx]
        b, c, h, w = x.shape
        if self.augment:
            xp = torch.nn.functional.pad(x.float(), (4, 4, 4, 4), mode='reflect')
            dx = torch.randint(0, 9, (b,), generator=self.gen)
            dy = torch.randint(0, 9, (b,), generator=self.gen)
            flip = torch.rand(b, generator=self.gen) < 0.5
            ar = torch.arange(w).view(1, 1, w)
            cols = dx.view(b, 1, 1) + torch.where(flip.view(b, 1, 1), (w - 1 - ar), ar)       # [b,1,w]
            rows = dy.view(b, 1, 1) + torch.arange(h).view(1, h, 1)                            # [b,h,1]
            flat = (rows * (w + 8) + cols).view(b, 1, h * w).expand(b, c, h * w)
            xf = xp.view(b, c, -1).gather(2, flat).view(b, c, h, w)
        else:
            xf = x.float()
        xf = xf.mul_(1.0 / 255.0).sub_(self.mean).mul_(self.inv_std)
        out = self._out_buffers(b)
        if out is None:
            xf = xf.to(self.out_dtype)
            if self.channels_last:
                xf = xf.contiguous(memory_format=torch.channels_last)
            return xf, y
        out[0].copy_(xf)
        out[1].copy_(y)
        return out

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.gen) if self.shuffle else torch.arange(n)
        nb = len(self)
        self._retire(0)                                # batches still held from the previous epoch
        if self.prefetch <= 0 or nb <= 1:
            for i in range(nb):
                yield self._track(self.make_batch(order[i * self.batch_size:(i + 1) * self.batch_size]))
        else:
            # background producer: batches are assembled (GIL released inside the C++ pass) while the step runs
            import queue
            import threading
            q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
            stop = threading.Event()

            def produce():
                try:
                    for i in range(nb):
                        if stop.is_set():
                            return
                        q.put(self.make_batch(order[i * self.batch_size:(i + 1) * self.batch_size]))
                    q.put(None)
                except Exception as e:  # noqa: BLE001
                    q.put(e)
            th = threading.Thread(target=produce, daemon=True)
            th.start()
            try:
                while True:
                    item = q.get()
                    if item is None:
                        break
                    if isinstance(item, Exception):
                        raise item
                    yield self._track(item)
            finally:
                stop.set()
                while not q.empty():
                    left = q.get_nowait()
                    if isinstance(left, tuple) and self.pin:       # assembled but never consumed: straight back to the pool
                        e = self._entry_of.pop(id(left[0]), None)
                        if e is not None:
                            self._free.put(e)
        self.dataset.epochs_completed += 1

    # reference API ---------------------------------------------------------------------------------------
    def next_batch(self, batch_size: Optional[int] = None):
        """TF-style pull API of the reference's custom loader (``my_data_loader.py:318-319``)."""
        if batch_size is not None and batch_size != self.batch_size:
            self.batch_size, self._it = batch_size, None
        it = getattr(self, "_it", None)
        if it is None:
            it = self._it = iter(self)
        try:
            return next(it)
        except StopIteration:
            self._it = iter(self)
            return next(self._it)


DataLoader = BatchLoader  # name used by the reference's single_machine.py


def mark_h2d_issued(host_tensor: torch.Tensor, stream=None) -> None:
    """Consumer side of the pinned-buffer protocol: call right after enqueueing the asynchronous host->device copy that reads
    ``host_tensor`` (a batch yielded by a pinning ``BatchLoader``).  Records a CUDA event on that stream; the loader waits on it
    before it lets the producer thread overwrite the buffer."""
    if host_tensor.is_cuda or not host_tensor.is_pinned():
        return
    ev = torch.cuda.Event()
    ev.record(stream if stream is not None else torch.cuda.current_stream())
    host_tensor._psb_evt = ev


def prepare_data(args, rank: int = 0, pin: bool = False, out_dtype=torch.float32, channels_last=False,
                 synthetic_train: int = 0, synthetic_test: int = 0) -> Tuple[BatchLoader, BatchLoader]:
    name = args.dataset
    if name not in _DIRS:
        raise ValueError("unknown dataset %r" % name)
    root = getattr(args, "data_root", ".")
    synth = getattr(args, "synthetic", None)
    train = test = None
    if not synth:
        train, test = _load_real(name, root, True), _load_real(name, root, False)
        if (train is None or test is None) and synth is False:
            raise FileNotFoundError("dataset %s not found under %s and --synthetic=false" % (name, root))
    if train is None or test is None:
        big = name == "ImageNet"
        train = synthetic_dataset(name, synthetic_train or (2048 if big else 16384), seed=0)
        test = synthetic_dataset(name, synthetic_test or (512 if big else 2048), seed=1)
    aug = name in ("Cifar10", "Cifar100", "SVHN")
    seed = int(getattr(args, "seed", 1)) * 1000 + rank
    train_loader = BatchLoader(train, args.batch_size, shuffle=True, augment=aug, seed=seed, pin=pin,
                               out_dtype=out_dtype, channels_last=channels_last, drop_last=True if len(train) >= args.batch_size else False)
    test_bs = getattr(args, "test_batch_size", None) or getattr(args, "eval_batch_size", 1000)
    test_loader = BatchLoader(test, min(test_bs, len(test)), shuffle=False, augment=False, seed=seed + 7)
    return train_loader, test_loader
