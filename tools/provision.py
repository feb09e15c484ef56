#!/usr/bin/env python
"""Cloud provisioning verbs — the part of the reference's ``tools/pytorch_ec2.py`` that talks to the AWS API (spot / on-demand
launch ``/root/reference/tools/pytorch_ec2.py:176-205``, the wait-until-running loops ``:209-256``, ``shutdown`` / ``clean_launch``
in the verb table ``:938-951``).  ``tools/cluster.py`` does everything that only needs ssh; this tool creates and destroys the
machines and hands them over by writing ``tools/inventory``.

The target here is one NVSwitch box (a ``p6-b200.48xlarge`` has the 8 B200s this framework is written for), so the default
plan is ONE instance; ``n_nodes > 1`` makes several independent boxes (or a gloo-tier job across them, see ``cluster.py launch``).

Verbs
  plan                  print the API requests ``launch`` would make (no credentials, no network)
  launch                request the instances (spot or on-demand), wait until they run and pass status checks, write the inventory
  clean_launch          ``shutdown`` whatever carries this cluster's tag, then ``launch``
  list                  instances tagged with this cluster name: id, state, type, public / private address
  wait                  block until every tagged instance is running (and, with --ssh, accepts ssh)
  write_inventory       (re)write ``tools/inventory`` from the running tagged instances
  shutdown              terminate the tagged instances and cancel open spot requests

Providers: ``ec2`` (boto3, imported lazily — it is not part of this image: install it on the operator machine) and ``dryrun``
(records the requests it is given and fabricates instances; used by ``plan`` and by the tests).  The configuration is a JSON
file (``--config``); values may reference other keys with ``%(key)s`` like the reference's self-interpolating ``Cfg``
(``pytorch_ec2.py:12-20``) — resolved once at load time instead of on every lookup.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import sys
import time
from typing import Dict, List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))

DEFAULTS: Dict[str, object] = {
    "name": "psb200",                       # cluster tag: every verb only ever touches instances tagged psb200-cluster=<name>
    "region": "us-east-1",
    "availability_zone": "",                # empty: let EC2 choose
    "instance_type": "p6-b200.48xlarge",    # 8 x B200 on one NVSwitch
    "n_nodes": 1,
    "method": "on-demand",                  # or "spot"
    "spot_price": "",                       # spot only; empty = on-demand price cap
    "image_id": "",                         # an AMI with the CUDA 12.9 driver stack + this repository's python environment
    "key_name": "",
    "security_group_ids": [],
    "subnet_id": "",
    "ssh_username": "ubuntu",
    "path_to_keyfile": "~/.ssh/%(key_name)s.pem",
    "volume_gb": 500,
    "shared_dir": "/mnt/psb200_shared",     # where cluster.py setup_nfs mounts the checkpoint directory
    "train_dir": "%(shared_dir)s/%(name)s",
}
TAG = "psb200-cluster"


def load_config(path: Optional[str], overrides: Optional[Dict[str, object]] = None) -> Dict[str, object]:
    cfg = dict(DEFAULTS)
    if path:
        with open(path) as f:
            cfg.update(json.load(f))
    cfg.update({k: v for k, v in (overrides or {}).items() if v is not None})
    # resolve %(key)s references to a fixed point (a reference cycle is a configuration error, not an endless loop)
    for _ in range(8):
        changed = False
        for k, v in list(cfg.items()):
            if isinstance(v, str) and "%(" in v:
                new = v % cfg
                changed |= new != v
                cfg[k] = new
        if not changed:
            break
    else:
        raise ValueError("configuration keys reference each other in a cycle")
    if cfg["method"] not in ("spot", "on-demand"):
        raise ValueError("method must be 'spot' or 'on-demand'")
    if int(cfg["n_nodes"]) < 1:
        raise ValueError("n_nodes must be >= 1")
    return cfg


def launch_request(cfg) -> Dict[str, object]:
    """The RunInstances request for this configuration (spot = the same request with market options: one code path, one request shape)."""
    req: Dict[str, object] = {
        "ImageId": cfg["image_id"], "InstanceType": cfg["instance_type"], "MinCount": int(cfg["n_nodes"]), "MaxCount": int(cfg["n_nodes"]),
        "KeyName": cfg["key_name"],
        "BlockDeviceMappings": [{"DeviceName": "/dev/sda1", "Ebs": {"VolumeSize": int(cfg["volume_gb"]), "VolumeType": "gp3", "DeleteOnTermination": True}}],
        "TagSpecifications": [{"ResourceType": "instance", "Tags": [{"Key": TAG, "Value": cfg["name"]}, {"Key": "Name", "Value": "%s-node" % cfg["name"]}]}],
    }
    if cfg["security_group_ids"]:
        req["SecurityGroupIds"] = list(cfg["security_group_ids"])
    if cfg["subnet_id"]:
        req["SubnetId"] = cfg["subnet_id"]
    if cfg["availability_zone"]:
        req["Placement"] = {"AvailabilityZone": cfg["availability_zone"]}
    if cfg["method"] == "spot":
        spot = {"SpotInstanceType": "one-time", "InstanceInterruptionBehavior": "terminate"}
        if cfg["spot_price"]:
            spot["MaxPrice"] = str(cfg["spot_price"])
        req["InstanceMarketOptions"] = {"MarketType": "spot", "SpotOptions": spot}
    return req


class Provider:
    """What a cloud has to offer this tool.  Instances are plain dicts: id, state, type, public_ip, private_ip."""

    def run_instances(self, request) -> List[dict]: raise NotImplementedError
    def describe(self, cluster: str) -> List[dict]: raise NotImplementedError
    def status_ok(self, ids: List[str]) -> bool: raise NotImplementedError
    def terminate(self, ids: List[str]) -> None: raise NotImplementedError
    def cancel_spot_requests(self, cluster: str) -> int: return 0


class DryRunProvider(Provider):
    """No network: remembers the requests, fabricates instances that become 'running' after ``boot_polls`` describe calls."""

    def __init__(self, boot_polls: int = 1):
        self.requests: List[dict] = []
        self.instances: List[dict] = []
        self.boot_polls = boot_polls
        self.terminated: List[str] = []

    def run_instances(self, request):
        self.requests.append(request)
        cluster = [t["Value"] for t in request["TagSpecifications"][0]["Tags"] if t["Key"] == TAG][0]
        new = []
        for _ in range(request["MaxCount"]):
            k = len(self.instances)
            inst = {"id": "i-dry%05d" % k, "state": "pending", "type": request["InstanceType"], "public_ip": "198.51.100.%d" % (10 + k),
                    "private_ip": "10.0.0.%d" % (10 + k), "cluster": cluster, "_polls": 0}
            self.instances.append(inst)
            new.append(inst)
        return [self._public(i) for i in new]

    @staticmethod
    def _public(i):
        return {k: v for k, v in i.items() if not k.startswith("_") and k != "cluster"}

    def describe(self, cluster):
        out = []
        for i in self.instances:
            if i["cluster"] != cluster or i["state"] == "terminated":
                continue
            if i["state"] == "pending":
                i["_polls"] += 1
                if i["_polls"] >= self.boot_polls:
                    i["state"] = "running"
            out.append(self._public(i))
        return out

    def status_ok(self, ids):
        return all(i["state"] == "running" for i in self.instances if i["id"] in ids)

    def terminate(self, ids):
        for i in self.instances:
            if i["id"] in ids:
                i["state"] = "terminated"
                self.terminated.append(i["id"])


class Ec2Provider(Provider):
    def __init__(self, region: str):
        try:
            import boto3                                  # noqa: PLC0415  (operator-side dependency, absent from the training image)
        except ImportError as e:
            raise SystemExit("the ec2 provider needs boto3 on this machine (pip install boto3); use `plan` to see the requests without it") from e
        self.client = boto3.client("ec2", region_name=region)

    @staticmethod
    def _inst(d):
        return {"id": d["InstanceId"], "state": d["State"]["Name"], "type": d["InstanceType"], "public_ip": d.get("PublicIpAddress", ""),
                "private_ip": d.get("PrivateIpAddress", "")}

    def run_instances(self, request):
        return [self._inst(d) for d in self.client.run_instances(**request)["Instances"]]

    def describe(self, cluster):
        flt = [{"Name": "tag:" + TAG, "Values": [cluster]}, {"Name": "instance-state-name", "Values": ["pending", "running", "stopping", "stopped"]}]
        out = []
        for page in self.client.get_paginator("describe_instances").paginate(Filters=flt):
            for r in page["Reservations"]:
                out += [self._inst(d) for d in r["Instances"]]
        return out

    def status_ok(self, ids):
        st = self.client.describe_instance_status(InstanceIds=ids)["InstanceStatuses"]
        return len(st) == len(ids) and all(s["InstanceStatus"]["Status"] == "ok" and s["SystemStatus"]["Status"] == "ok" for s in st)

    def terminate(self, ids):
        if ids:
            self.client.terminate_instances(InstanceIds=ids)

    def cancel_spot_requests(self, cluster):
        reqs = self.client.describe_spot_instance_requests(Filters=[{"Name": "tag:" + TAG, "Values": [cluster]}, {"Name": "state", "Values": ["open", "active"]}])
        ids = [r["SpotInstanceRequestId"] for r in reqs["SpotInstanceRequests"]]
        if ids:
            self.client.cancel_spot_instance_requests(SpotInstanceRequestIds=ids)
        return len(ids)


# ---- verbs --------------------------------------------------------------------------------------------------------------
def wait_running(p: Provider, cfg, timeout_s: float = 1800.0, poll_s: float = 5.0, check_ssh: bool = False, sleep=time.sleep) -> List[dict]:
    """Block until all n_nodes tagged instances are running and pass the provider's status checks (reference :209-256: two polling loops,
    here one with a deadline instead of forever)."""
    deadline = time.time() + timeout_s
    while True:
        inst = p.describe(cfg["name"])
        running = [i for i in inst if i["state"] == "running"]
        if len(running) >= int(cfg["n_nodes"]) and p.status_ok([i["id"] for i in running]):
            if not check_ssh or all(_ssh_open(i["public_ip"] or i["private_ip"]) for i in running):
                return running
        if time.time() > deadline:
            raise TimeoutError("%d of %d instances running after %.0f s" % (len(running), int(cfg["n_nodes"]), timeout_s))
        sleep(poll_s)


def _ssh_open(addr: str, port: int = 22, timeout: float = 3.0) -> bool:
    try:
        with socket.create_connection((addr, port), timeout=timeout):
            return True
    except OSError:
        return False


def write_inventory(instances: List[dict], cfg, path: str) -> None:
    """``cluster.py``'s inventory format: <address> [alias] [ssh_user] [ssh_key]; the first node is the head (NFS export, gloo master)."""
    key = os.path.expanduser(str(cfg["path_to_keyfile"]))
    with open(path, "w") as f:
        f.write("# written by tools/provision.py for cluster %s\n" % cfg["name"])
        for k, i in enumerate(sorted(instances, key=lambda d: d["id"])):
            f.write("%s %s %s %s\n" % (i["public_ip"] or i["private_ip"], "node%d" % k, cfg["ssh_username"], key))


def shutdown(p: Provider, cfg) -> List[str]:
    ids = [i["id"] for i in p.describe(cfg["name"])]
    p.terminate(ids)
    p.cancel_spot_requests(cfg["name"])
    return ids


def launch(p: Provider, cfg, inventory: str, wait_timeout_s: float = 1800.0, poll_s: float = 5.0, sleep=time.sleep) -> List[dict]:
    existing = p.describe(cfg["name"])
    if existing:
        raise RuntimeError("cluster %r already has %d instance(s): use clean_launch or shutdown first" % (cfg["name"], len(existing)))
    for k in ("image_id", "key_name"):
        if not cfg[k]:
            raise ValueError("configuration key %r is required to launch" % k)
    p.run_instances(launch_request(cfg))
    running = wait_running(p, cfg, wait_timeout_s, poll_s, sleep=sleep)
    write_inventory(running, cfg, inventory)
    return running


def main(argv=None, provider: Optional[Provider] = None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("verb", choices=["plan", "launch", "clean_launch", "list", "wait", "write_inventory", "shutdown"])
    ap.add_argument("--config", default=None, help="JSON file; missing keys take the defaults printed by `plan`")
    ap.add_argument("--provider", default="ec2", choices=["ec2", "dryrun"])
    ap.add_argument("--inventory", default=os.path.join(HERE, "inventory"))
    ap.add_argument("--name"); ap.add_argument("--n-nodes", type=int); ap.add_argument("--method", choices=["spot", "on-demand"])
    ap.add_argument("--timeout", type=float, default=1800.0)
    ap.add_argument("--ssh", action="store_true", help="wait: also require port 22 to accept connections")
    a = ap.parse_args(argv)
    cfg = load_config(a.config, {"name": a.name, "n_nodes": a.n_nodes, "method": a.method})
    if a.verb == "plan":
        print(json.dumps({"config": cfg, "run_instances": launch_request(cfg)}, indent=1))
        return 0
    p = provider or (DryRunProvider() if a.provider == "dryrun" else Ec2Provider(str(cfg["region"])))
    if a.verb in ("launch", "clean_launch"):
        if a.verb == "clean_launch":
            gone = shutdown(p, cfg)
            print("terminated %d instance(s)" % len(gone))
        running = launch(p, cfg, a.inventory, a.timeout)
        print("running: " + ", ".join("%s (%s)" % (i["id"], i["public_ip"] or i["private_ip"]) for i in running))
        print("inventory written to %s — next: tools/cluster.py sync / setup_nfs / launch" % a.inventory)
    elif a.verb == "list":
        for i in p.describe(cfg["name"]):
            print("%-20s %-10s %-18s %-16s %s" % (i["id"], i["state"], i["type"], i["public_ip"], i["private_ip"]))
    elif a.verb == "wait":
        print("%d running" % len(wait_running(p, cfg, a.timeout, check_ssh=a.ssh)))
    elif a.verb == "write_inventory":
        write_inventory([i for i in p.describe(cfg["name"]) if i["state"] == "running"], cfg, a.inventory)
    elif a.verb == "shutdown":
        print("terminated: " + (", ".join(shutdown(p, cfg)) or "nothing"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
