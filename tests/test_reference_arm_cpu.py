"""The reference arm: the unmodified reference runs end to end on the mpi4py/blosc stand-ins (CPU, LeNet, 2 workers)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))


def _have_reference():
    return os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "src", "distributed_nn.py")) or os.path.isdir("/root/reference/src")


@pytest.mark.skipif(not _have_reference(), reason="reference tree not present (baseline/install_ref.sh)")
def test_reference_arm_lenet_cpu():
    import ref_arm
    out = ref_arm.run(2, 4, 1, network="LeNet", dataset="MNIST", batch_size=64, use_gpu=False, timeout_s=300.0)
    assert "unavailable" not in out, out
    assert out["impl"] == "reference" and out["value"] > 0 and out["config"]["global_batch"] == 128
    assert out["final_loss"] < out["first_loss"]          # the PS update reaches the workers: the loss moves


def test_refshim_waitany_orders_by_arrival(tmp_path):
    """Two processes: rank 1 sends tags in a scrambled order, rank 0's waitany returns them in arrival order."""
    import subprocess
    code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "baseline", "refshim"))
from mpi4py import MPI
import numpy as np
comm = MPI.COMM_WORLD
if comm.Get_rank() == 0:
    reqs = [comm.irecv(source=1, tag=88 + i) for i in range(6)]
    raw = np.zeros(4); rr = comm.Irecv([raw, MPI.DOUBLE], source=1, tag=5)
    got = []
    for _ in range(6):
        st = MPI.Status()
        i, msg = MPI.Request.waitany(requests=reqs, status=st)
        assert st.tag == 88 + i and st.source == 1 and msg == ("m", i)
        got.append(i)
    rr.wait()
    assert got == [3, 1, 5, 0, 2, 4], got
    assert raw.tolist() == [1.0, 2.0, 3.0, 4.0]
    assert comm.bcast("w", root=0) == "w"
    print("ok")
else:
    import time
    for i in [3, 1, 5, 0, 2, 4]:
        comm.isend(("m", i), dest=0, tag=88 + i).wait(); time.sleep(0.01)
    comm.Isend([np.array([1.0, 2.0, 3.0, 4.0]), MPI.DOUBLE], dest=0, tag=5)
    assert comm.bcast(None, root=0) == "w"
''' % ROOT
    procs = []
    for r in range(2):
        env = dict(os.environ, REFSHIM_RANK=str(r), REFSHIM_SIZE="2", REFSHIM_DIR=str(tmp_path / "sock"))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=60)[0] for p in procs]
    assert procs[0].returncode == 0 and procs[1].returncode == 0, outs
    assert "ok" in outs[0]
