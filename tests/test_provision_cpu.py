"""Provisioning verbs (tools/provision.py) against the dry-run provider: request shape, wait loop, inventory hand-over to cluster.py,
shutdown — the API-facing part of the reference's ``tools/pytorch_ec2.py:176-256,938-951`` without a network."""
import importlib.util
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


prov = _load("provision")


def _cfg(tmp_path, **kw):
    base = {"name": "t1", "image_id": "ami-123", "key_name": "k", "n_nodes": 2}
    base.update(kw)
    p = tmp_path / "c.json"
    p.write_text(json.dumps(base))
    return prov.load_config(str(p))


def test_config_interpolation_and_validation(tmp_path):
    cfg = _cfg(tmp_path, shared_dir="/nfs", name="exp7")
    assert cfg["train_dir"] == "/nfs/exp7" and cfg["path_to_keyfile"].endswith("k.pem")
    with pytest.raises(ValueError):
        _cfg(tmp_path, method="reserved")
    with pytest.raises(ValueError):
        _cfg(tmp_path, a="%(b)s", b="%(a)s x")


def test_spot_and_on_demand_requests(tmp_path):
    od = prov.launch_request(_cfg(tmp_path))
    assert od["MinCount"] == od["MaxCount"] == 2 and "InstanceMarketOptions" not in od
    assert {"Key": prov.TAG, "Value": "t1"} in od["TagSpecifications"][0]["Tags"]
    sp = prov.launch_request(_cfg(tmp_path, method="spot", spot_price="12.5", availability_zone="us-east-1b"))
    assert sp["InstanceMarketOptions"]["SpotOptions"]["MaxPrice"] == "12.5" and sp["Placement"] == {"AvailabilityZone": "us-east-1b"}


def test_launch_waits_writes_inventory_and_shutdown_terminates(tmp_path):
    cfg = _cfg(tmp_path)
    p = prov.DryRunProvider(boot_polls=3)
    naps = []
    inv = str(tmp_path / "inventory")
    running = prov.launch(p, cfg, inv, poll_s=0.0, sleep=naps.append)
    assert len(running) == 2 and len(naps) >= 1, "the instances boot after a few polls: launch must have waited"
    cluster = _load("cluster")
    nodes = cluster.read_inventory(inv)
    assert [n["alias"] for n in nodes] == ["node0", "node1"] and nodes[0]["user"] == "ubuntu" and nodes[0]["addr"].startswith("198.51.100.")
    with pytest.raises(RuntimeError):
        prov.launch(p, cfg, inv)                       # a second launch of the same cluster is refused
    assert sorted(prov.shutdown(p, cfg)) == sorted(i["id"] for i in running)
    assert p.describe("t1") == []


def test_wait_times_out_instead_of_looping_forever(tmp_path):
    cfg = _cfg(tmp_path)
    p = prov.DryRunProvider(boot_polls=10 ** 9)
    p.run_instances(prov.launch_request(cfg))
    with pytest.raises(TimeoutError):
        prov.wait_running(p, cfg, timeout_s=0.0, poll_s=0.0, sleep=lambda s: None)


def test_cli_plan_and_clean_launch(tmp_path, capsys):
    c = tmp_path / "c.json"
    c.write_text(json.dumps({"name": "cli", "image_id": "ami-1", "key_name": "k"}))
    assert prov.main(["plan", "--config", str(c)]) == 0
    plan = json.loads(capsys.readouterr().out)
    assert plan["run_instances"]["InstanceType"] == "p6-b200.48xlarge" and plan["config"]["n_nodes"] == 1
    p = prov.DryRunProvider()
    inv = str(tmp_path / "inv")
    assert prov.main(["clean_launch", "--config", str(c), "--inventory", inv], provider=p) == 0
    assert prov.main(["clean_launch", "--config", str(c), "--inventory", inv], provider=p) == 0
    assert len(p.terminated) == 1 and len(p.describe("cli")) == 1
