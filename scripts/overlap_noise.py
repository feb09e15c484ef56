"""Noise floor of test_nvl_overlap_with_graph_matches_plain: the same 2-GPU job run twice in the plain configuration, once with --overlap (no graph) and
once with --overlap + graph; prints the pairwise cosine of the weight updates.  plain-vs-plain is what bf16 + fp32-atomic summation order alone does."""
import itertools
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(out, extra, env_extra=None):
    cmd = [sys.executable, "-m", "ps_pytorch_b200.launch", "-n", "2", "--timeout", "200", os.path.join(ROOT, "tests", "mp_job.py"), "--out", out,
           "--transport", "nvl", "--device", "cuda", "--dtype", "bf16", "--steps", "4", "--network", "ResNet18", "--bs", sys.argv[1] if len(sys.argv) > 1 else "32",
           "--colocate", "1", "--backend", "sm100"] + extra
    env = dict(os.environ, PYTHONPATH=ROOT)
    env.update(env_extra or {})
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=260)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return torch.load(out)


from test_gpu_multi import _init  # noqa: E402
init = _init("ResNet18")
d = tempfile.mkdtemp()
cfgs = {"plain1": (["--graph", "0", "--overlap", "0"], None), "plain2": (["--graph", "0", "--overlap", "0"], None),
        "plain_nosink": (["--graph", "0", "--overlap", "0"], {"PSB200_GRAD_SINKS": "0", "PSB200_TC_WGRAD_STREAM": "0"}),
        "overlap": (["--graph", "0", "--overlap", "1"], None), "overlap_graph": (["--graph", "1", "--overlap", "1"], None)}
upd = {}
for k, (extra, env) in cfgs.items():
    w = run(os.path.join(d, k + ".pt"), extra, env)["w"]
    upd[k] = torch.cat([(x - i).flatten() for x, i in zip(w, init)])
for a, b in itertools.combinations(upd, 2):
    print("%-14s %-14s cos %.4f  norm ratio %.4f" % (a, b, torch.nn.functional.cosine_similarity(upd[a], upd[b], dim=0).item(), (upd[a].norm() / upd[b].norm()).item()))
