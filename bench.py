#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): ResNet-18 / CIFAR-10-shaped synthetic data, bf16, per-worker batch 1024
(the reference's ``run_pytorch.sh`` configuration), synchronous parameter-server training — images/s over the whole
job, device-timed, max over ranks.

    python bench.py --gpus 1 --steps 30 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 8 ...

``--impl reference`` is the reference arm: the UNMODIFIED hwang595/ps_pytorch tree (verbatim copy in the git-ignored
``baseline/_ref``) run through its own entry point ``src/distributed_nn.py`` with the flags of ``src/run_pytorch.sh``
(``baseline/ref_arm.py``; mpirun / mpi4py / blosc, none of which exist in this image, are replaced by the stand-ins in
``baseline/refshim``).  ``--impl nccl-baseline`` runs ``baseline/nccl_ps.py``: the reference algorithm re-expressed with
NCCL reduce/broadcast + stock torch.nn (the operative baseline of BASELINE.md §2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _tc_conv():
    try:
        from ps_pytorch_b200.ops import conv as C
        return C.tc_conv_enabled()
    except Exception:
        return False


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference", "nccl-baseline"])
    ap.add_argument("--network", type=str, default="ResNet18")
    ap.add_argument("--dataset", type=str, default="Cifar10")
    ap.add_argument("--batch-size", type=int, default=1024, help="per-worker batch (reference run_pytorch.sh: 1024)")
    ap.add_argument("--topology", type=str, default="ps", choices=["ps", "sharded"])
    ap.add_argument("--grad-dtype", type=str, default="bf16", choices=["bf16", "fp8", "fp32"])
    ap.add_argument("--backend", type=str, default="auto", choices=["auto", "sm100", "library"])
    ap.add_argument("--no-colocate", action="store_true", help="reference topology: the PS GPU runs no worker")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--overlap", action="store_true")
    ap.add_argument("--optimizer", type=str, default="sgd")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--no-stage", dest="stage", action="store_false", help="do not prefetch the next batch's H2D copy")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: fix the GLOBAL batch (per-worker batch = global / workers)")
    ap.add_argument("--no-baseline", action="store_true", help="skip the in-job NCCL-baseline measurement")
    ap.add_argument("--num-aggregate", type=int, default=0, help="K of P (default: all workers)")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.rows, self.proc = [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """Keep only the samples taken inside the timed region."""
        self.rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _t, r in self.rows[-max(1, len(self.rows) // 4):]]

    def stop(self, ngpu):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clk, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            if isinstance(r, tuple):
                r = r[1]
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9 or not f[0].isdigit() or int(f[0]) >= ngpu:
                continue
            try:
                clk.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(clk) if clk else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(clk), "reasons": sorted(reasons)}


def main():
    a = parse()
    if a.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:          # under torchrun every rank gets here: rank 0 launches the whole reference job
            return 0
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import ref_arm
        sampler = ClockSampler()
        sampler.start()
        out = ref_arm.run(a.gpus, a.steps, a.warmup, network=a.network, dataset=a.dataset, batch_size=a.batch_size)
        clocks = sampler.stop(a.gpus)
        if "unavailable" not in out:
            out["clocks"] = clocks
        print(json.dumps(out))
        return 0
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:    # convenience: self-launch
            from ps_pytorch_b200.launch import launch
            return launch(a.gpus, [sys.executable, os.path.abspath(__file__)] + sys.argv[1:])
        raise SystemExit("WORLD_SIZE=%d but --gpus %d" % (world, a.gpus))
    assert torch.cuda.is_available(), "bench.py needs a GPU"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    if a.impl == "nccl-baseline":
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import nccl_ps
        return nccl_ps.bench(a, rank, local, world)

    from ps_pytorch_b200 import ops
    from ps_pytorch_b200.engine import PSJob
    from ps_pytorch_b200.config import input_shape_for, num_classes_for
    colocate = not a.no_colocate
    nw_expected = world if (colocate or world == 1 or a.topology == "sharded") else world - 1
    if a.global_batch:
        assert a.global_batch % nw_expected == 0, "--global-batch must divide by the number of workers"
        a.batch_size = a.global_batch // nw_expected
    job = PSJob(network=a.network, dataset=a.dataset, batch_size=a.batch_size, device=dev, dtype="bf16", lr=0.1, momentum=0.9,
                optimizer=a.optimizer, topology=a.topology, colocate=colocate, grad_dtype=a.grad_dtype, overlap=a.overlap,
                cuda_graph=not a.no_graph, backend=a.backend, num_aggregate=a.num_aggregate or None)
    nworkers = job.comm.num_workers
    c, h, w = input_shape_for(a.dataset)
    B = a.batch_size
    ncls = num_classes_for(a.dataset)
    # Synthetic data of the named shape, resident in PINNED host memory; 8 distinct batches rotate so no step re-reads the previous
    # step's input (and the ~3 GB/step activation working set is >> the 126 MB L2 anyway).  The images are CLASS-SEPARABLE (a fixed
    # random template per class, shared by all ranks, plus unit noise), so the loss has to fall well below ln(#classes): a dead
    # broadcast, a skipped reduce or diverging replicas would show up in first_loss/final_loss and in the replica check below.
    nb = 8
    tg = torch.Generator().manual_seed(7)
    templates = torch.randn(ncls, c, h, w, generator=tg)
    g = torch.Generator().manual_seed(1234 + rank)
    ys = [torch.randint(0, ncls, (B,), generator=g).pin_memory() for _ in range(nb)]
    xs = [(0.7 * templates[y] + torch.randn(B, c, h, w, generator=g)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).pin_memory() for y in ys]
    h2d = xs[0].numel() * xs[0].element_size() + ys[0].numel() * 8 if job.is_worker else 0

    def run(nsteps, read_loss):
        last = None
        for i in range(nsteps):
            last = job.step(xs[i % nb], ys[i % nb])
            if a.stage:
                job.stage(xs[(i + 1) % nb], ys[(i + 1) % nb])     # next batch's H2D overlaps this step
            if read_loss and last is not None:
                last = float(last)            # D2H read of the step's result, every step
        return last

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(nsteps, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        last = run(nsteps, read_loss)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), last

    sampler = ClockSampler()
    if rank == 0:
        sampler.start()                   # nvidia-smi needs ~1 s to start streaming: launch it before the warm-up
    first = run(1, True)                  # step 1 (also the CUDA-graph capture): the loss of the untrained net, ~ln(#classes)
    run(max(a.warmup, 3) - 1, False)      # rest of the untimed warm-up
    barrier()
    if hasattr(job.comm, "check_error"):
        job.comm.check_error()
    l0 = ops.launch_count
    tw0 = time.perf_counter()
    ms, wall_ms, loss = timed(a.steps, False)
    tw1 = time.perf_counter()
    launches = ops.launch_count - l0
    clocks = None
    if ms < 350.0:                        # too short for a 100 ms sampler: keep the load on (untimed, all ranks) so 3+ samples exist
        run(int(500.0 / max(ms / a.steps, 0.1)) + 1, False)
        torch.cuda.synchronize()
        tw1 = time.perf_counter()
    if rank == 0:
        time.sleep(0.15)
        sampler.window(tw0, tw1 + 0.1)
        clocks = sampler.stop(world)
    e2e = None
    if not a.skip_e2e:
        ems, ewall, loss = timed(a.steps, True)
        e2e_ms = max(ems, ewall)          # the D2H read blocks the host: wall time is the honest end-to-end clock
        e2e = {"value": nworkers * B * a.steps / (e2e_ms * 1e-3), "unit": "images/s", "ms_per_step": e2e_ms / a.steps,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 if job.is_worker else 0}
    if hasattr(job.comm, "check_error"):
        job.comm.check_error()
    # ---- exposed communication, measured directly: CUDA events around wait-weights / push / reduce+update+broadcast (engine.PSJob.time_comm)
    exposed = exposed_max = None
    if hasattr(job, "time_comm"):
        job.time_comm(True)
        loss2 = run(min(a.steps, 20), False)
        barrier()
        cm = job.comm_ms_per_step()
        job.time_comm(False)
        # Every rank has the same step period T = compute_r + C_r (synchronous training), where C_r is the bracketed time (spins included).
        # The rank with the longest compute has the smallest C_r: MIN over ranks = the communication the critical path is exposed to;
        # MAX additionally contains the time faster ranks spend waiting for the slowest one (skew, not communication).
        tmin = torch.tensor([cm if cm is not None else 1e9], dtype=torch.float64)
        tmax = torch.tensor([cm if cm is not None else 0.0], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        exposed, exposed_max = float(tmin[0]), float(tmax[0])
        loss = loss2 if loss2 is not None else loss
    # ---- self-verification: every replica must hold exactly the bf16 rounding of rank 0's fp32 master, and the loss must have fallen
    barrier()
    if hasattr(job.comm, "gather_master") and job.master is not None:
        job.comm.gather_master(job.master.master)          # sharded topology: rank 0's master is only authoritative for its own shard
    check = {}
    if job.worker is not None:
        rep = job.worker.compute.weights.flat
        csum = int(rep.view(torch.int16).to(torch.int64).sum().item())
        cabs = float(rep.float().abs().sum().item())
    else:
        csum, cabs = None, None
    sums = [None] * world
    if world > 1:
        dist.all_gather_object(sums, (csum, cabs))
    else:
        sums = [(csum, cabs)]
    if rank == 0:
        m = job.master
        want = m.master.flat.to(torch.bfloat16)
        wsum = int(want.view(torch.int16).to(torch.int64).sum().item())
        have = [s_[0] for s_ in sums if s_[0] is not None]
        check = {"replicas": len(have), "replicas_identical": len(set(have)) == 1, "replicas_equal_bf16_of_master": all(v == wsum for v in have),
                 "weight_abs_sum": sums[-1][1]}
    lt = torch.tensor([float(loss) if loss is not None else float("nan")])
    lf = float(first) if first is not None else float("nan")
    if world > 1:                             # a pure-PS rank 0 (--no-colocate) has no loss of its own: take worker 1's
        both = [None] * world
        dist.all_gather_object(both, (lf, float(lt[0])))
        good = [b for b in both if b[1] == b[1]]
        lf, lt = good[0][0], torch.tensor([good[0][1]])
    transport_name = job.comm.name + ":" + getattr(getattr(job.comm, "arena", None), "backend", "-")
    nccl_base = None
    if not a.no_baseline and a.network == "ResNet18" and a.dataset == "Cifar10" and not a.global_batch:
        try:
            job.close()
            job = None
            sys.path.insert(0, os.path.join(ROOT, "baseline"))
            import nccl_ps
            nccl_base = nccl_ps.run(a, rank, local, world, use_graph=(world == 1))
        except Exception as e:  # noqa: BLE001
            nccl_base = {"error": repr(e)[:200]}
    if rank == 0:
        value = nworkers * B * a.steps / (ms * 1e-3)
        out = {"metric": "ResNet-18/CIFAR-10 images/sec (whole box, device-timed, max over ranks)" if a.network == "ResNet18" else a.network + " images/sec",
               "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
               "higher_is_better": True, "scaling": "strong" if a.global_batch else "weak", "vs_baseline": None, "dtype": "bf16",
               "data": "synthetic (class-separable images of the dataset shape: per-class template + noise; random-init weights)",
               "impl": "ours",
               "config": {"model": a.network, "dataset_shape": "%s %dx%dx%d" % (a.dataset, c, h, w), "global_batch": nworkers * B,
                          "per_worker_batch": B, "seq_len": None,
                          "parallelism": "%s: 1 server + %d workers on %d GPUs%s" % (a.topology, nworkers, world, " (worker co-located on the PS GPU)" if colocate else ""),
                          "grad_wire": a.grad_dtype, "optimizer": a.optimizer + " momentum 0.9, fused in the PS kernel", "backend": a.backend,
                          "transport": transport_name,
                          "cuda_graph": not a.no_graph,
                          "conv": "tcgen05 implicit-GEMM (ours)" if _tc_conv() else "cuDNN (library); all other layers + PS kernels ours",
                          "l2": "inputs rotate over 8 distinct pinned host batches; per-step activation working set (~3 GB) >> 126 MB L2"},
               "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "wall_ms_per_step": wall_ms / a.steps,
               "first_loss": lf, "final_loss": float(lt[0]), "loss_fell": bool(float(lt[0]) < 0.8 * lf) if lf == lf else None,
               "replica_check": check, "exposed_comm_ms": exposed, "comm_incl_peer_wait_ms_max": exposed_max,
               "nccl_baseline": nccl_base,
               "vs_nccl_baseline": (value / nccl_base["value"]) if (nccl_base and nccl_base.get("value")) else None}
        print(json.dumps(out))
    if job is not None:
        job.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main() or 0)
