"""Observability: phase timers (CUDA events on GPU, perf_counter on CPU), NVTX ranges, the reference's
worker log line and a JSONL sink.

The reference reads ``time.time()`` around un-synchronised GPU work, so its Forward/Backward
numbers are launch times.  ``PhaseTimer`` records CUDA events on the launching stream and resolves them after
the step, so per-phase numbers are device times and cost no host sync inside the step.
"""
from __future__ import annotations

import contextlib
import json
import logging
import re
import time
from typing import Dict, Optional

import torch

logging.basicConfig()
logger = logging.getLogger("ps_pytorch_b200")
logger.setLevel(logging.INFO)

WORKER_LOG_FORMAT = ('Worker: {}, Step: {}, Epoch: {} [{}/{} ({:.0f}%)], Loss: {:.4f}, Time Cost: {:.4f}, '
                     'FetchWeight: {:.4f}, Forward: {:.4f}, Backward: {:.4f}, Comm Cost: {:.4f}')
# superset fields appended after the reference's (``distributed_worker.py:169-173``)
WORKER_LOG_EXTRA = ', Msg(MB): {:.3f}, Img/s: {:.1f}'
WORKER_LOG_RE = re.compile(
    r'Worker: (?P<rank>\d+), Step: (?P<step>\d+), Epoch: (?P<epoch>\d+) \[(?P<seen>\d+)/(?P<total>\d+) \((?P<pct>[\d.]+)%\)\], '
    r'Loss: (?P<loss>[-\d.naninf]+), Time Cost: (?P<time>[\d.]+), FetchWeight: (?P<fetch>[\d.]+), Forward: (?P<fwd>[\d.]+), '
    r'Backward: (?P<bwd>[\d.]+), Comm Cost: (?P<comm>[\d.]+)(?:, Msg\(MB\): (?P<msg_mb>[\d.]+), Img/s: (?P<ips>[\d.]+))?')


class PhaseTimer:
    """``with timer.phase("forward"): ...`` — durations in seconds, resolved lazily."""

    def __init__(self, device):
        self.cuda = torch.device(device).type == "cuda"
        self._open: Dict[str, tuple] = {}
        self._done: Dict[str, float] = {}

    @contextlib.contextmanager
    def phase(self, name: str):
        if self.cuda:
            torch.cuda.nvtx.range_push(name)
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record()
            try:
                yield
            finally:
                b.record()
                torch.cuda.nvtx.range_pop()
                self._open[name] = (a, b)
        else:
            t0 = time.perf_counter()
            try:
                yield
            finally:
                self._done[name] = self._done.get(name, 0.0) + time.perf_counter() - t0

    def discard(self) -> None:
        self._open.clear()
        self._done.clear()

    def resolve(self) -> Dict[str, float]:
        if self.cuda and self._open:
            torch.cuda.synchronize()
            for k, (a, b) in self._open.items():
                self._done[k] = self._done.get(k, 0.0) + a.elapsed_time(b) * 1e-3
            self._open.clear()
        out, self._done = self._done, {}
        return out


class JsonlSink:
    def __init__(self, path: Optional[str]):
        self.f = open(path, "a") if path else None

    def write(self, **kw):
        if self.f:
            self.f.write(json.dumps(kw) + "\n")
            self.f.flush()

    def close(self):
        if self.f:
            self.f.close()


class MetricsExporter:
    """Prometheus endpoint (``--metrics-port P`` → rank r serves http://0.0.0.0:P+r/metrics): the fields of the worker log line as gauges, so a
    long job can be scraped instead of grepped.  A no-op when the port is 0 or ``prometheus_client`` is missing."""

    def __init__(self, port: int, rank: int):
        self.on = False
        if not port:
            return
        try:
            from prometheus_client import CollectorRegistry, Counter, Gauge, start_http_server
        except ImportError:
            logger.warning("--metrics-port given but prometheus_client is not installed")
            return
        self.reg = CollectorRegistry()
        lab = {"rank": str(rank)}
        mk = lambda n, d: Gauge("psb200_" + n, d, list(lab), registry=self.reg).labels(**lab)
        self.step = mk("step", "last finished training step")
        self.loss = mk("loss", "loss of the last logged step")
        self.step_s = mk("step_seconds", "wall time of the last logged step")
        self.ips = mk("images_per_second", "per-worker throughput of the last logged step")
        self.fwd = mk("forward_seconds", "device time of the forward pass")
        self.bwd = mk("backward_seconds", "device time of the backward pass")
        self.comm = mk("comm_seconds", "time in the gradient push")
        self.msg = Counter("psb200_pushed_megabytes", "gradient payload pushed", list(lab), registry=self.reg).labels(**lab)
        start_http_server(port + rank, registry=self.reg)
        self.on = True

    def update(self, step, loss, dur, ips, fwd, bwd, comm, msg_mb):
        if not self.on:
            return
        self.step.set(step); self.loss.set(loss); self.step_s.set(dur); self.ips.set(ips)
        self.fwd.set(fwd); self.bwd.set(bwd); self.comm.set(comm); self.msg.inc(max(0.0, msg_mb or 0.0))


def format_worker_line(rank, step, epoch, seen, total, loss, t_iter, t_fetch, t_fwd, t_bwd, t_comm, msg_mb=None, ips=None):
    pct = 100.0 * seen / max(total, 1)
    s = WORKER_LOG_FORMAT.format(rank, step, epoch, seen, total, pct, loss, t_iter, t_fetch, t_fwd, t_bwd, t_comm)
    if msg_mb is not None:
        s += WORKER_LOG_EXTRA.format(msg_mb, ips or 0.0)
    return s


def parse_worker_line(line: str):
    m = WORKER_LOG_RE.search(line)
    if not m:
        return None
    d = m.groupdict()
    out = {k: (int(v) if k in ("rank", "step", "epoch", "seen", "total") else (float(v) if v is not None else None))
           for k, v in d.items()}
    return out
