#!/usr/bin/env bash
# Generic GPU job wrapper: `bash scripts/gpu_call.sh <name> <command...>` runs the command with PYTHONPATH set, a log under gpurun_out/<name>.log,
# and prints its tail.  (Always run `python __graft_entry__.py build` locally before gpurun: the in-tree .so is what travels.)
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
name=$1; shift
"$@" > gpurun_out/$name.log 2>&1
echo "rc=$?"; tail -${TAIL:-40} gpurun_out/$name.log
