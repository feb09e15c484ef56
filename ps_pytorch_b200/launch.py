"""``python -m ps_pytorch_b200.launch -n N script.py --flags`` — the in-house replacement for
``mpirun -n N --hostfile … python distributed_nn.py`` (``/root/reference/src/run_pytorch.sh:1``).

Spawns N processes on this node (one per GPU on a B200 box, or N CPU ranks for the gloo tier) with the
torchrun-style environment (RANK, LOCAL_RANK, WORLD_SIZE, MASTER_ADDR=127.0.0.1, MASTER_PORT) and supervises them:
if any rank dies the rest of the process *group started here* is terminated (exact PIDs, never by pattern), so a
failed rank cannot leave spinning peers behind — the failure detection the reference lacks.
``torch.distributed.run`` is an equally valid launcher; the scripts only read the environment.

``--max-restarts R`` adds fail-stop recovery: when a rank dies the job is torn down and started again (new rendezvous port,
``PSB200_RESTART_COUNT`` = attempt number, ``--restart-args`` appended — by default ``--resume``, so the PS reloads the newest
``model_step_<N>`` + optimizer sidecar and the step counter continues from there), at most R times.
"""
from __future__ import annotations

import argparse
import os
import signal
import socket
import subprocess
import sys
import time


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(nprocs: int, cmd, master_port: int = 0, env_extra=None, timeout: float = 0.0, log_dir: str = "") -> int:
    port = master_port or _free_port()
    procs = []
    for r in range(nprocs):
        env = dict(os.environ)
        env.update({"RANK": str(r), "LOCAL_RANK": str(r), "WORLD_SIZE": str(nprocs), "LOCAL_WORLD_SIZE": str(nprocs),
                    "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
        env.update(env_extra or {})
        out = open(os.path.join(log_dir, "rank%d.log" % r), "w") if log_dir else None
        procs.append(subprocess.Popen(cmd, env=env, stdout=out, stderr=subprocess.STDOUT if out else None))
    t0, rc = time.time(), 0
    try:
        alive = set(range(nprocs))
        while alive:
            for r in list(alive):
                c = procs[r].poll()
                if c is not None:
                    alive.discard(r)
                    if c != 0:
                        rc = rc or c
                        sys.stderr.write("[launch] rank %d exited with %d; terminating peers\n" % (r, c))
                        for q in alive:
                            procs[q].terminate()
            if timeout and time.time() - t0 > timeout:
                sys.stderr.write("[launch] timeout after %.0fs; terminating\n" % timeout)
                rc = rc or 124
                for q in alive:
                    procs[q].terminate()
                timeout = 0
            time.sleep(0.05)
    except KeyboardInterrupt:
        for p in procs:
            p.send_signal(signal.SIGINT)
        rc = 130
    for p in procs:
        try:
            p.wait(timeout=10)
        except subprocess.TimeoutExpired:
            p.kill()
    return rc


def launch_with_restarts(nprocs: int, cmd, restart_args, max_restarts: int, master_port: int = 0, timeout: float = 0.0, log_dir: str = "") -> int:
    """Run ``cmd`` on ``nprocs`` ranks; on a failed attempt (not a timeout, not Ctrl-C) start over with ``restart_args`` appended."""
    attempt = 0
    while True:
        rc = launch(nprocs, cmd + (list(restart_args) if attempt else []), 0 if attempt else master_port,
                    env_extra={"PSB200_RESTART_COUNT": str(attempt)}, timeout=timeout, log_dir=log_dir)
        if rc in (0, 124, 130) or attempt >= max_restarts:
            return rc
        attempt += 1
        sys.stderr.write("[launch] attempt failed with %d; restart %d/%d with %s\n" % (rc, attempt, max_restarts, " ".join(restart_args) or "(no extra args)"))
        time.sleep(1.0)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-n", "--nprocs", type=int, required=True)
    ap.add_argument("--master-port", type=int, default=0)
    ap.add_argument("--timeout", type=float, default=0.0)
    ap.add_argument("--log-dir", type=str, default="")
    ap.add_argument("--max-restarts", type=int, default=0, help="relaunch the whole job up to this many times when a rank dies")
    ap.add_argument("--restart-args", type=str, default="--resume", help="arguments appended to the script on restarts")
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    cmd = [sys.executable, a.script] + a.script_args
    if a.max_restarts > 0:
        sys.exit(launch_with_restarts(a.nprocs, cmd, a.restart_args.split(), a.max_restarts, a.master_port, a.timeout, a.log_dir))
    sys.exit(launch(a.nprocs, cmd, a.master_port, timeout=a.timeout, log_dir=a.log_dir))


if __name__ == "__main__":
    main()
