#!/usr/bin/env bash
# One-time offline "install" of the reference into the git-ignored baseline/_ref (travels to the GPU box with gpurun).
# pip cannot install it (no setup.py / pyproject.toml), so the tree is copied verbatim; nothing in it is edited.
set -u
cd "$(dirname "$0")/.."
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference 2>&1 | tail -2
if [ ! -f baseline/_ref/src/distributed_nn.py ]; then
  mkdir -p baseline/_ref
  cp -r /root/reference/. baseline/_ref/
  rm -rf baseline/_ref/.git
fi
ls baseline/_ref/src | head -30
