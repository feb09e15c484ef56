#!/usr/bin/env bash
# Node bootstrap — the role of the reference's tools/pre_run.sh (conda + pytorch + python-blosc + mpi4py installs).
# Nothing is installed here: the framework needs only PyTorch and the CUDA toolkit that are already on a B200 box.
# This script checks them and builds the in-tree sm_100a extensions.
set -euo pipefail
cd "$(dirname "$0")/.."
python - <<'PY'
import shutil, sys, torch
print("python", sys.version.split()[0], "| torch", torch.__version__, "| cuda available:", torch.cuda.is_available(), "| gpus:", torch.cuda.device_count())
assert shutil.which("nvcc"), "nvcc not on PATH (needed to build the sm_100a kernels)"
import __graft_entry__ as g
g.build()
print("extensions built:", "ok")
PY
