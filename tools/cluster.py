#!/usr/bin/env python
"""Cluster operations — the role of the reference's ``tools/pytorch_ec2.py`` (boto3 + paramiko EC2 manager,
``/root/reference/tools/pytorch_ec2.py:938-964``) for machines you already have: an inventory file instead of the AWS API.

Verbs (same vocabulary as the reference where it makes sense):

  get_hosts              write ``hosts`` / ``hosts_alias`` / ``hosts_address`` from the inventory (reference :656-708)
  run_command  "<cmd>"   run a shell command on every node in parallel over ssh (reference :854-877)
  status                 per node: GPUs, utilisation, running python jobs (the reference's idle detection :304-336)
  sync                   rsync this repository to every node (reference ``update_git_dir.sh`` / ``remote_script.sh``)
  setup_nfs              export <dir> from the first node and mount it on the others (reference :880-900)
  launch  -- <args>      start ``src/distributed_nn.py`` on every node: one ``ps_pytorch_b200.launch`` per node is NOT how the
                         single-box engine scales (NVLink tier = one NVSwitch box); this starts independent single-box jobs,
                         or a gloo-tier multi-node job with --gloo (RANK offsets per node)
  kill_all               terminate the jobs started by ``launch`` (exact PIDs from the pid files it wrote — never by pattern;
                         the reference's ``killall python`` / ``kill_all_python`` :821-852)

Inventory (``tools/inventory``): one node per line ``<address> [alias] [ssh_user] [ssh_key]``; ``localhost`` runs locally.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shlex
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PID_DIR = "/tmp/psb200_pids"


def read_inventory(path):
    nodes = []
    with open(path) as f:
        for line in f:
            line = line.split("#")[0].strip()
            if not line:
                continue
            p = line.split()
            nodes.append({"addr": p[0], "alias": p[1] if len(p) > 1 else "node%d" % len(nodes),
                          "user": p[2] if len(p) > 2 else None, "key": p[3] if len(p) > 3 else None})
    return nodes


def ssh_cmd(node, cmd):
    if node["addr"] in ("localhost", "127.0.0.1"):
        return ["bash", "-lc", cmd]
    base = ["ssh", "-o", "StrictHostKeyChecking=no", "-o", "BatchMode=yes"]
    if node["key"]:
        base += ["-i", node["key"]]
    target = (node["user"] + "@" if node["user"] else "") + node["addr"]
    return base + [target, cmd]


def run_on(node, cmd, timeout=600):
    try:
        r = subprocess.run(ssh_cmd(node, cmd), capture_output=True, text=True, timeout=timeout)
        return node, r.returncode, r.stdout, r.stderr
    except subprocess.TimeoutExpired:
        return node, 124, "", "timeout"


def parallel(nodes, cmd, timeout=600):
    with cf.ThreadPoolExecutor(max_workers=max(1, len(nodes))) as ex:
        return list(ex.map(lambda n: run_on(n, cmd, timeout), nodes))


def show(results):
    rc = 0
    for node, code, out, err in results:
        print("=== %s (%s) rc=%d" % (node["alias"], node["addr"], code))
        if out.strip():
            print(out.rstrip())
        if err.strip():
            print(err.rstrip(), file=sys.stderr)
        rc = rc or code
    return rc


def get_hosts(nodes, outdir):
    with open(os.path.join(outdir, "hosts"), "w") as f:
        for n in nodes:
            f.write("%s\t%s\n" % (n["addr"], n["alias"]))
    with open(os.path.join(outdir, "hosts_alias"), "w") as f:
        f.write("\n".join(n["alias"] for n in nodes) + "\n")
    with open(os.path.join(outdir, "hosts_address"), "w") as f:
        f.write("\n".join(n["addr"] for n in nodes) + "\n")
    print("wrote hosts, hosts_alias, hosts_address (%d nodes) to %s" % (len(nodes), outdir))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("verb", choices=["get_hosts", "run_command", "status", "sync", "setup_nfs", "launch", "kill_all"])
    ap.add_argument("--inventory", default=os.path.join(HERE, "inventory"))
    ap.add_argument("--remote-dir", default="~/ps_pytorch_b200")
    ap.add_argument("--nfs-dir", default="/mnt/psb200_shared")
    ap.add_argument("--gloo", action="store_true", help="launch: one multi-node gloo-tier job instead of independent single-box jobs")
    ap.add_argument("--nproc-per-node", type=int, default=0)
    ap.add_argument("--out-dir", default=HERE, help="get_hosts: where to write hosts / hosts_alias / hosts_address")
    ap.add_argument("rest", nargs="*", help="run_command / launch: the command or flags (put them after --)")
    argv = sys.argv[1:]
    tail = []
    if "--" in argv:                                   # options may come before or after the verb; everything after -- is passed through
        i = argv.index("--")
        argv, tail = argv[:i], argv[i + 1:]
    a = ap.parse_args(argv)
    nodes = read_inventory(a.inventory)
    rest = list(a.rest) + tail
    if a.verb == "get_hosts":
        return get_hosts(nodes, a.out_dir)
    if a.verb == "run_command":
        return show(parallel(nodes, " ".join(rest)))
    if a.verb == "status":
        cmd = ("nvidia-smi --query-gpu=index,name,utilization.gpu,memory.used --format=csv,noheader 2>/dev/null || echo 'no GPU'; "
               "for f in %s/*.pid; do [ -f \"$f\" ] && (kill -0 $(cat $f) 2>/dev/null && echo running $f || echo stale $f); done" % PID_DIR)
        return show(parallel(nodes, cmd, 60))
    if a.verb == "sync":
        rc = 0
        for n in nodes:
            if n["addr"] in ("localhost", "127.0.0.1"):
                continue
            target = (n["user"] + "@" if n["user"] else "") + n["addr"] + ":" + a.remote_dir
            ssh = "ssh -o StrictHostKeyChecking=no" + (" -i " + n["key"] if n["key"] else "")
            rc = rc or subprocess.call(["rsync", "-az", "--delete", "--exclude", ".git", "--exclude", "gpurun_out", "-e", ssh, ROOT + "/", target])
        return rc
    if a.verb == "setup_nfs":
        head = nodes[0]
        show([run_on(head, "sudo mkdir -p {d} && sudo chmod 777 {d} && echo '{d} *(rw,sync,no_subtree_check)' | sudo tee -a /etc/exports && sudo exportfs -ra".format(d=a.nfs_dir))])
        return show(parallel(nodes[1:], "sudo mkdir -p {d} && sudo mount -t nfs {h}:{d} {d}".format(d=a.nfs_dir, h=head["addr"])))
    if a.verb == "launch":
        results = []
        world_nodes = len(nodes)
        for i, n in enumerate(nodes):
            nproc = a.nproc_per_node or 8
            if a.gloo:
                env = "MASTER_ADDR=%s MASTER_PORT=29511 NODE_RANK=%d NNODES=%d" % (nodes[0]["addr"], i, world_nodes)
                inner = "python -m torch.distributed.run --nnodes=%d --node-rank=%d --nproc-per-node=%d --master-addr %s --master-port 29511 src/distributed_nn.py --transport gloo %s" % (
                    world_nodes, i, nproc, nodes[0]["addr"], " ".join(shlex.quote(x) for x in rest))
            else:
                env = ""
                inner = "python -m ps_pytorch_b200.launch -n %d src/distributed_nn.py %s" % (nproc, " ".join(shlex.quote(x) for x in rest))
            cmd = "mkdir -p {p} && cd {d} && ({e} nohup {c} > job_{a}.log 2>&1 & echo $! > {p}/job_{a}.pid) && echo started".format(
                p=PID_DIR, d=a.remote_dir, e=env, c=inner, a=n["alias"])
            results.append(run_on(n, cmd, 60))
        return show(results)
    if a.verb == "kill_all":
        cmd = "for f in %s/*.pid; do [ -f \"$f\" ] && kill -- -$(ps -o pgid= -p $(cat $f) | tr -d ' ') 2>/dev/null; rm -f $f; done; echo done" % PID_DIR
        return show(parallel(nodes, cmd, 60))


if __name__ == "__main__":
    sys.exit(main() or 0)
