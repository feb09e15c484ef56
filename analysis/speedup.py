"""Log analysis (replaces the reference's two notebooks, ``analysis/Speedup_Comparisons_LeNet.ipynb`` and
``analysis/Speedups_with_GradCompression.ipynb``): parse worker log lines, take per step the max ("normal") and the
min ("ideal") ``Time Cost`` over workers, sum over the first N steps, and print speed-ups against the smallest run.

    python analysis/speedup.py --steps 100 1:logs/w1.log 2:logs/w2.log 4:logs/w4.log ...
"""
from __future__ import annotations

import argparse
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200.utils.obs import parse_worker_line  # noqa: E402


def per_step_times(path):
    by_step = defaultdict(dict)
    with open(path) as f:
        for line in f:
            d = parse_worker_line(line)
            if d:
                by_step[d["step"]][d["rank"]] = d["time"]
    return by_step


def totals(by_step, nsteps, skip=1):
    steps = sorted(s for s in by_step if s > skip)[:nsteps]
    normal = sum(max(by_step[s].values()) for s in steps)
    ideal = sum(min(by_step[s].values()) for s in steps)
    return normal, ideal, len(steps)


def per_worker_mean(by_step, skip=1):
    acc, cnt = defaultdict(float), defaultdict(int)
    for s, d in by_step.items():
        if s <= skip:
            continue
        for r, t in d.items():
            acc[r] += t
            cnt[r] += 1
    return {r: acc[r] / cnt[r] for r in acc}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("runs", nargs="+", help="<num_workers>:<logfile>")
    a = ap.parse_args()
    rows = []
    for item in a.runs:
        n, path = item.split(":", 1)
        normal, ideal, used = totals(per_step_times(path), a.steps)
        rows.append((int(n), normal, ideal, used, per_worker_mean(per_step_times(path))))
    rows.sort()
    base_n, base_normal = rows[0][0], rows[0][1]
    print("%8s %12s %12s %10s %10s %6s" % ("workers", "normal_s", "ideal_s", "speedup", "ideal_sp", "steps"))
    for n, normal, ideal, used, pw in rows:
        print("%8d %12.4f %12.4f %10.2f %10.2f %6d" % (n, normal, ideal, base_n * base_normal / normal if normal else 0,
                                                     base_n * base_normal / ideal if ideal else 0, used))
    n, _, _, _, pw = rows[-1]
    if pw:
        print("per-worker mean iteration time at %d workers: fastest %.4f s, slowest %.4f s" % (n, min(pw.values()), max(pw.values())))


if __name__ == "__main__":
    main()
