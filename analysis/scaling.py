"""Scaling table from ``bench.py`` JSON lines — the device-timed counterpart of the log-derived speed-up curves in the reference's notebooks
(``/root/reference/analysis/Speedup_Comparisons_LeNet.ipynb``): images/s, speed-up and efficiency against the smallest run, exposed communication and the
in-job NCCL baseline, one row per GPU count.

    python analysis/scaling.py profiles/run29_bench_n1.json profiles/run27_n2_default.json profiles/run27_n4_default.json profiles/run27_n8_default.json
"""
from __future__ import annotations

import json
import sys


def load(path):
    with open(path) as f:
        text = f.read().strip()
    return json.loads(text.splitlines()[-1])


def table(paths):
    rows = sorted((load(p) for p in paths), key=lambda d: d["n_gpus"])
    base = rows[0]
    out = ["| GPUs | scaling | ms/step | images/s | speed-up | efficiency | exposed comm ms | NCCL baseline images/s | ours / NCCL |", "|---|---|---|---|---|---|---|---|---|"]
    for d in rows:
        n = d["n_gpus"] / base["n_gpus"]
        sp = d["value"] / base["value"]
        # weak scaling: ideal = n x the base rate; strong scaling (fixed global batch): the same definition applies to images/s
        nb = (d.get("nccl_baseline") or {}).get("value")
        out.append("| %d | %s | %.3f | %.0f | %.2f | %.3f | %s | %s | %s |" % (
            d["n_gpus"], d.get("scaling", "weak"), d["ms_per_step"], d["value"], sp, sp / n,
            "%.3f" % d["exposed_comm_ms"] if d.get("exposed_comm_ms") is not None else "-",
            "%.0f" % nb if nb else "-", "%.2f" % (d["value"] / nb) if nb else "-"))
    return "\n".join(out)


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    print(table(sys.argv[1:]))
